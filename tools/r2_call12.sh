#!/bin/bash
# Round 2, GPU call 12: why does the view's tile-kernel hash differ from linon_compile_check's?  (module text + options of both)
set -u
OUT=gpurun_out/r2c12
mkdir -p "$OUT"
LINON_DUMP_SOURCE="$OUT/view_src.cu" timeout 200 python tools/render_once.py --frames 2 > "$OUT/render_once.json" 2>&1
LINON_CHECK_QUIET_ZONE=2 LINON_DUMP_SOURCE="$OUT/check_src.cu" LINON_DUMP_CUBIN_HASH="$OUT/check_hash.txt" timeout 200 python -c "
from linon_b200.main_view import PredefinedFunction, compile_check
from linon_b200 import _capi as capi
compile_check(PredefinedFunction.MirageSphericalSigmoid.to_code(), flags=1, math_mode=capi.MATH_FAST)
" > "$OUT/check.log" 2>&1
head -c 1500 "$OUT/view_src.cu" | head -3 | cut -c1-1200; echo; head -3 "$OUT/check_src.cu" | cut -c1-1200; echo; cat "$OUT/check_hash.txt"; tail -1 "$OUT/render_once.json" | cut -c1-300
cmp "$OUT/view_src.cu" "$OUT/check_src.cu" && echo SAME_SOURCE
