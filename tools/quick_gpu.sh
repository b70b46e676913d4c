#!/usr/bin/env bash
# Fast kernel-iteration call (about a minute of box time):
#   gpurun --timeout 600 -- 'bash tools/quick_gpu.sh TAG'
# runs the rasterizer parity tests, then tools/variants.py at 24 and 3 views for the default library and every tagged build.
TAG="${1:-q}"
OUT=gpurun_out
mkdir -p "$OUT"
(timeout 400 python -m pytest tests/test_rasterizer_gpu.py -m gpu -x -q 2>&1 | tail -5) > "$OUT/${TAG}_pytest.log"
timeout 200 python tools/variants.py 24 > "$OUT/${TAG}_variants24.txt" 2>&1
timeout 200 python tools/variants.py 3 > "$OUT/${TAG}_variants3.txt" 2>&1
cat "$OUT/${TAG}_pytest.log" "$OUT/${TAG}_variants24.txt" "$OUT/${TAG}_variants3.txt"
