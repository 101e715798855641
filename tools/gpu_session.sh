#!/bin/bash
# developer helper: one gpurun call = GPU tests + table/kernel timings of every library variant
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
: > gpurun_out/table_perf.jsonl
python tools/table_perf.py --check >> gpurun_out/table_perf.jsonl 2>gpurun_out/table_perf.err
NOA_DCS_TABLE_LAUNCH=combined python tools/table_perf.py --check >> gpurun_out/table_perf.jsonl 2>>gpurun_out/table_perf.err
shopt -s nullglob
for lib in variants/*.so; do
  NOA_DCS_LIB=$PWD/$lib python tools/table_perf.py >> gpurun_out/table_perf.jsonl 2>>gpurun_out/table_perf.err
done
cat gpurun_out/table_perf.jsonl
