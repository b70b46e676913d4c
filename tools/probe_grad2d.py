"""Stage-level accuracy of the backward blend: the CUDA path's per-(view, Gaussian) 2-D gradient records (grad2d in the workspace)
against the oracle's double-precision accumulators (acc2d), component by component.  Run once per library build
(TOPO4D_B200_LIB=...) to compare reduction variants.
    python tools/probe_grad2d.py [generic|topo4d] [blend_px]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import parity  # noqa: E402
from topo4d_b200 import synth  # noqa: E402

opacity = sys.argv[1] if len(sys.argv) > 1 else "generic"
px = int(sys.argv[2]) if len(sys.argv) > 2 else 2
scene = synth.head_scene(60000, seed=0, sh_degree=3, opacity=opacity)
cams = [synth.ring_cameras(24)[7]]
H, W = 1080, 1920
rng = np.random.default_rng(1)
gC = rng.normal(size=(1, 3, H, W)).astype(np.float32)
gD = rng.normal(size=(1, 1, H, W)).astype(np.float32)
gA = rng.normal(size=(1, 1, H, W)).astype(np.float32)
ref = parity.oracle_forward(scene, cams, H, W, 3, (0, 0, 0))
acc = ref[0]["state"].backward(gC[0], gD[0], gA[0], return_acc2d=True)["acc2d"]          # [N,10] float64
out = parity.cuda_forward(scene, cams, H, W, 3, (0, 0, 0), blend_px=px)
parity.cuda_backward(out, gC, gD, gA)
g2 = out["state"].view()["grad2d"][0].cpu().numpy().astype(np.float64)                    # [N,12]
# grad2d record: (dpix.x, dpix.y, dconA, dconB | dconC, dopacity, ddepth, _ | dr, dg, db, _)
ours = np.stack([g2[:, 0], g2[:, 1], g2[:, 2], g2[:, 3], g2[:, 4], g2[:, 5], g2[:, 8], g2[:, 9], g2[:, 10], g2[:, 6]], 1)
names = ["dpix.x", "dpix.y", "dconA", "dconB", "dconC", "dopacity", "dr", "dg", "db", "ddepth"]
print(os.environ.get("TOPO4D_B200_LIB", "default library"), opacity, "px", px)
for i, n in enumerate(names):
    a, b = ours[:, i], acc[:, i]
    scale = np.abs(b).max()
    rel = np.abs(a - b) / (np.abs(b) + 1e-3 * scale)
    print(f"  {n:9s} max|ref| {scale:10.3e}  max relerr {rel.max():.3e}  mean relerr {rel.mean():.3e}  bias {np.sum(a - b) / (np.sum(np.abs(b)) + 1e-300):+.2e}")
