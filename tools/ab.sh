#!/bin/bash
# kernel A/B: tools/ab.sh <variant>...  (variants built by tools/build_variant.sh; "prod" = the in-tree library)
cd "$(dirname "$0")/.."
for v in "$@"; do
  if [ "$v" = prod ]; then unset PAGK_LIB; else export PAGK_LIB=$PWD/tools/libpagk_$v.so; fi
  timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu 2>gpurun_out/ab_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v: features/s %.4g  feat-iter/s %.4g  ms/step %.4f  serial ms/step %.4f  lk_ms %.4f  e2e %.4g'%(d['value'], d['feature_iterations_per_sec'], d['ms_per_step'], d['serial']['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))" || tail -5 gpurun_out/ab_$v.err
done
