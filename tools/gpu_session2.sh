#!/bin/bash
# developer helper (2-GPU box): all GPU tests incl. the multi-GPU ones, then bench N=1 and N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s14_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/s14_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/s14_pytest_gpu.log
tail -15 gpurun_out/s14_pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s14_bench_n1.json 2> gpurun_out/s14_bench_n1.err; echo "bench n1 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s14_bench_n2.json 2> gpurun_out/s14_bench_n2.err; echo "bench n2 exit $?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/s14_bench_ref.json 2> gpurun_out/s14_bench_ref.err; echo "ref exit $?"
tail -c 1500 gpurun_out/s14_bench_n1.err; tail -c 1500 gpurun_out/s14_bench_n2.err
head -c 3000 gpurun_out/s14_bench_n1.json; echo; head -c 1500 gpurun_out/s14_bench_n2.json
