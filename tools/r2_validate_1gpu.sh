#!/bin/bash
# full 1-GPU validation + round-2 profile captures
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu.log
tail -5 gpurun_out/r2/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2/bench_E.json 2> gpurun_out/r2/bench_E.err; tail -c 600 gpurun_out/r2/bench_E.json
timeout 300 python bench.py --workload B > gpurun_out/r2/bench_B.json 2> gpurun_out/r2/bench_B.err
timeout 300 python bench.py --workload C > gpurun_out/r2/bench_C.json 2> gpurun_out/r2/bench_C.err
timeout 300 python bench.py --workload D > gpurun_out/r2/bench_D.json 2> gpurun_out/r2/bench_D.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2/launches_E.csv python bench.py --steps 1 --warmup 1 --profile --no-extra-legs --no-cpu-baseline > gpurun_out/r2/ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_assign_tc' -c 6 -o gpurun_out/r2/assign_tc_E python bench.py --steps 1 --warmup 1 --profile --no-extra-legs --no-cpu-baseline > gpurun_out/r2/ncu_full_assign_tc.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_spmm' -c 3 -o gpurun_out/r2/spmm_E python bench.py --steps 1 --warmup 1 --profile --no-extra-legs --no-cpu-baseline > gpurun_out/r2/ncu_full_spmm.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_rs_scatter' -c 8 -o gpurun_out/r2/rs_scatter_E python bench.py --steps 1 --warmup 1 --profile --no-extra-legs --no-cpu-baseline > gpurun_out/r2/ncu_full_rs_scatter.log 2>&1
ls -la gpurun_out/r2
