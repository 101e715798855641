#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/s4_exchange_self.jsonl
for mode in split combined default; do
  NOA_DCS_TABLE_LAUNCH=$mode timeout 300 python tools/exchange_self_perf.py >> gpurun_out/s4_exchange_self.jsonl 2>gpurun_out/s4_err.log
done
cat gpurun_out/s4_exchange_self.jsonl; tail -3 gpurun_out/s4_err.log
