#!/bin/bash
# the GPU suite and the bench lines that go into profiles/ (run through gpurun; everything lands in gpurun_out/)
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_gpu_tests.log
python bench.py > gpurun_out/r02_bench_config_B.json 2> gpurun_out/r02_bench_config_B.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
python bench.py --config A --pairs 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_config_A_batch1.json 2> gpurun_out/r02_bench_config_A_batch1.err
python bench.py --config C --pairs 16 --steps 20 --warmup 3 > gpurun_out/r02_bench_config_C_batch16.json 2> gpurun_out/r02_bench_config_C_batch16.err
python bench.py --config E --pairs 4 --steps 20 --warmup 3 > gpurun_out/r02_bench_config_E_batch4.json 2> gpurun_out/r02_bench_config_E_batch4.err
tail -c 600 gpurun_out/r02_bench_config_B.err
