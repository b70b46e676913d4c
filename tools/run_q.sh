#!/usr/bin/env bash
# One quick GPU-box call while iterating on kernels:  gpurun --timeout 900 -- 'bash tools/run_q.sh TAG [all]'
# parity tests (all of -m gpu with "all", else the rasterizer file), then tools/variants.py at 24 and 3 views.
TAG="${1:-q}"
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-}" = "all" ]; then T=tests; else T=tests/test_rasterizer_gpu.py; fi
(timeout 500 python -m pytest $T -m gpu -x -q 2>&1 | tail -4) > $OUT/${TAG}_pytest.log
timeout 150 python tools/variants.py 24 > $OUT/${TAG}_variants24.txt 2>&1
timeout 150 python tools/variants.py 3 > $OUT/${TAG}_variants3.txt 2>&1
cat $OUT/${TAG}_pytest.log $OUT/${TAG}_variants24.txt $OUT/${TAG}_variants3.txt
