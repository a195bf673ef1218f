#!/bin/bash
# Round 2, GPU call 35: an edited shading block of ray_color on the device (five edits x four march kernels against frames of the
# reference's shader with the same edit), and the march-level constant edits again after the shade_hit refactor
set -u
OUT=gpurun_out/r2c35
mkdir -p "$OUT"
timeout 150 python -m pytest tests/test_gpu_units.py -q -m gpu -p no:cacheprovider -rs -k "shading_block or march_level_constants" > "$OUT/tests.log" 2>&1; echo "tests rc=$?"
tail -4 "$OUT/tests.log"
