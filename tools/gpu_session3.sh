#!/bin/bash
# developer helper (1 GPU): tests, ncu launch list of the bench command, ncu --set full of every
# kernel, then the bench itself (never under a profiler).  Usage: tools/gpu_session3.sh <tag>
tag=${1:-s3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
tail -6 gpurun_out/${tag}_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --builds-per-step 4 --no-extras > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --metrics sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum --clock-control none --import-source on \
    -k regex:'vmap_kernel|table_' -s 8 -c 8 -o gpurun_out/${tag}_prof -f python tools/profile_kernels.py > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench exit $?"
tail -c 600 gpurun_out/${tag}_bench_n1.err
head -c 1500 gpurun_out/${tag}_bench_n1.json
