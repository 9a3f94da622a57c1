#!/bin/bash
# quick GPU check used during kernel work: parity tests, then a short bench summary
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('features/s %.3g  feat-iter/s %.4g  ms/step %.3f  lk_ms %.3f  e2e %.3g'%(d['value'], d['feature_iterations_per_sec'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))"
