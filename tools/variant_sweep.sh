#!/bin/bash
shopt -s nullglob
# developer helper: run tools/quick_perf.py against every library build under variants/
for lib in noa_b200/libnoa_dcs_b200.so variants/*.so; do
  echo "== $lib"
  NOA_DCS_LIB=$PWD/$lib python tools/quick_perf.py 2>&1 | grep -v -E "table/|fp64_probe|\"A\"" | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('   %-16s %8.3f ms  %7.3f Gevals/s' % (d['kernel']+(str(d.get('min_points','')) ), d['best_ms'], d.get('Gevals_s',0)))
"
done
