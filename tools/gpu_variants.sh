#!/bin/bash
# developer helper: tools/table_perf.py against the in-tree library and every build under variants/
mkdir -p gpurun_out
: > gpurun_out/variants_perf.jsonl
for lib in noa_b200/libnoa_dcs_b200.so variants/*.so; do
  NOA_DCS_LIB=$PWD/$lib timeout 300 python tools/table_perf.py --check >> gpurun_out/variants_perf.jsonl 2>> gpurun_out/variants_err.log
done
python - <<'PY'
import json
for l in open('gpurun_out/variants_perf.jsonl'):
    d = json.loads(l)
    print(d['lib'].split('/')[-1], ' '.join(f"{k}={v:.4g}" for k, v in d.items() if isinstance(v, float) and ('table1000' in k or 'Gevals' in k)), 'mismatches', d.get('table_mismatches'))
PY
tail -3 gpurun_out/variants_err.log
