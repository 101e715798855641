#!/bin/bash
# developer helper: element-wise kernels (tools/stream_perf.py) for the in-tree library and every build under variants/
for lib in noa_b200/libnoa_dcs_b200.so variants/*.so; do
  echo "== $lib"
  NOA_DCS_LIB=$PWD/$lib python tools/stream_perf.py 2>&1 | grep -E "2\^24|pinned brems"
done
