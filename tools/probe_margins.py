import sys, json
sys.path.insert(0, "/root/repo")
from tests import parity
from tests.test_rasterizer_gpu import _cfg2, _CFG2_FLIPS
for op in ("topo4d", "generic"):
    scene, cams = _cfg2(op)
    for px in (2, 4, 1):
        m = parity.compare(scene, cams, 1080, 1920, 3, (0, 0, 0), noise_floor=True, blend_px=px, cache_key="m-" + op)
        print(op, px, {k: float("%.3g" % v) for k, v in m.items() if k.startswith("grad_relerr") or k.startswith("fp32_floor")}, m["n_contrib_mismatch"], flush=True)
