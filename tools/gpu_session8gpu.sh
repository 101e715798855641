#!/bin/bash
# developer helper (8-GPU box): every GPU test incl. the 2/4/8-rank ones, the exchange forms side by
# side (also for every library build under variants/), bench N=8 (full line), N=4, N=2, N=1
tag=${1:-s8g}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
tail -4 gpurun_out/${tag}_pytest_gpu.log
: > gpurun_out/${tag}_exchange.jsonl
for lib in noa_b200/libnoa_dcs_b200.so variants/*.so; do
  [ -f $lib ] || continue
  NOA_DCS_LIB=$PWD/$lib timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/table_exchange_bench.py 2>/dev/null | grep "^{" >> gpurun_out/${tag}_exchange.jsonl
done
cat gpurun_out/${tag}_exchange.jsonl
for n in 8 4 2; do
  extra=""; [ $n != 8 ] && extra="--no-extras"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 $extra > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err; echo "bench n$n exit $?"
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench n1 exit $?"
tail -c 400 gpurun_out/${tag}_bench_n8.err
for n in 1 2 4 8; do python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_bench_n$n.json').read().strip().splitlines()[-1])
print($n, d['value'], d['ms_per_build'], d['e2e']['value'], d['parity']['bit_exact'], d['parity'].get('every_rank_equals_single_gpu_build'))"; done
