OUT=gpurun_out
(timeout 300 python -m pytest tests/test_rasterizer_gpu.py -m gpu -x -q 2>&1 | tail -3) > $OUT/r02q2_pytest.log
timeout 120 python tools/variants.py 24 > $OUT/r02q2_variants24.txt 2>&1
timeout 120 python tools/variants.py 3 > $OUT/r02q2_variants3.txt 2>&1
timeout 200 python tools/probe_streams.py 3 3,1 > $OUT/r02q2_streams3.txt 2>&1
timeout 200 python tools/probe_streams.py 6 6,3,2,1 > $OUT/r02q2_streams6.txt 2>&1
timeout 200 python tools/probe_streams.py 12 12,6,4,3,2 > $OUT/r02q2_streams12.txt 2>&1
timeout 300 python tools/probe_streams.py 24 24,12,8,6,4,3 20 > $OUT/r02q2_streams24.txt 2>&1
cat $OUT/r02q2_*.txt $OUT/r02q2_pytest.log
