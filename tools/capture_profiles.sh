#!/bin/bash
# the ncu launch list, the full captures of K3b / K3a / K1 and the per-warp phase profile of one config-B bench step (run through gpurun)
set -x
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/pre.json 2> gpurun_out/pre.err || exit 1
# (ncu serialises the kernels: a launch that shares the device would run alone on its share, so the launch list is taken with every
# launch filling the device)
PAGK_BENCH_SHARE=1 PAGK_BENCH_E2E_SHARE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1
for k in lk_lanes lk_template pyramid_fused; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r02_$k python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 python tools/lanes_prof.py tools/libpagk_prof8.so > gpurun_out/r02_lanes_phase_profile.txt 2>&1
ls -la gpurun_out/*.ncu-rep
