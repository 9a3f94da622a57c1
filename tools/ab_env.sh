#!/bin/bash
# A/B of environment knobs on the in-tree library: tools/ab_env.sh "NAME=VAL ..." "NAME=VAL ..." (an empty string = defaults)
cd "$(dirname "$0")/.."
for v in "$@"; do
  env $v timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu 2>gpurun_out/ab_env.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$v] features/s %.4g  feat-iter/s %.4g  ms/step %.4f  serial ms/step %.4f  lk_ms %.4f  e2e %.4g  parity %s'%(d['value'], d['feature_iterations_per_sec'], d['ms_per_step'], d['serial']['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], (d.get('parity') or {}).get('bit_exact_all_outputs')))" || tail -5 gpurun_out/ab_env.err
done
