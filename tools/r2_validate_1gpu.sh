#!/bin/bash
# full 1-GPU validation + round-2 bench lines and launch list (run on the GPU box: bash tools/r2_validate_1gpu.sh)
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu.log
tail -3 gpurun_out/r2/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2/bench_E.json 2> gpurun_out/r2/bench_E.err
timeout 300 python bench.py --workload B > gpurun_out/r2/bench_B.json 2> gpurun_out/r2/bench_B.err
timeout 300 python bench.py --workload C > gpurun_out/r2/bench_C.json 2> gpurun_out/r2/bench_C.err
timeout 300 python bench.py --workload D > gpurun_out/r2/bench_D.json 2> gpurun_out/r2/bench_D.err
timeout 300 python tools/svd_bench.py > gpurun_out/r2/svd_C.json 2> gpurun_out/r2/svd_C.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2/launches_E.csv python bench.py --steps 1 --warmup 1 --profile --no-extra-legs --no-cpu-baseline > gpurun_out/r2/ncu_launch.log 2>&1
ls -la gpurun_out/r2
