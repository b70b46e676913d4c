"""Dev analysis (not a test): how many (region, Gaussian) pairs of the reference tile lists can touch any pixel of
the region, for regions 16x16 / 8x16 / 8x8 / 8x4.  CPU only, uses the oracle's forward geometry."""
import sys
import numpy as np
from oracle import gs_oracle
from topo4d_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
W, H = 1920, 1080
scene = synth.head_scene(n, seed=0, sh_degree=None, opacity="topo4d")
cam = synth.ring_cameras(24, w=W, h=H)[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
out = gs_oracle.forward(scene["means3D"], scene["opacities"], H, W, cam.tanfovx, cam.tanfovy, (0, 0, 0), cam.viewmatrix,
                        cam.projmatrix, cam.campos, colors_precomp=scene["colors_precomp"], scales=scene["scales"],
                        rotations=scene["rotations"])
st = out[4]
geo, b = st.geometry(), st.binning()
ids = b["ids"].astype(np.int64)
tiles = (b["keys"] >> np.uint64(32)).astype(np.int64)
gx = (W + 15) // 16
tx, ty = tiles % gx, tiles // gx
xy, co = geo["xy"][ids], geo["conic_opacity"][ids]
thr = -np.log(255.0 * np.maximum(co[:, 3], 1e-30))
I = ids.size
px = np.arange(16, dtype=np.float32)
alive = np.zeros((I, 16, 16), bool)
for s in range(0, I, 20000):
    e = min(I, s + 20000)
    dx = xy[s:e, 0, None] - (tx[s:e, None] * 16 + px[None])           # [n,16] x
    dy = xy[s:e, 1, None] - (ty[s:e, None] * 16 + px[None])           # [n,16] y
    A, B, C = co[s:e, 0, None, None], co[s:e, 1, None, None], co[s:e, 2, None, None]
    pw = -0.5 * (A * dx[:, None, :] ** 2 + C * dy[:, :, None] ** 2) - B * dx[:, None, :] * dy[:, :, None]
    alive[s:e] = (pw >= thr[s:e, None, None]) & (pw <= 0)
    # clip to the image
    X = tx[s:e, None] * 16 + np.arange(16)[None]; Y = ty[s:e, None] * 16 + np.arange(16)[None]
    alive[s:e] &= (X < W)[:, None, :] & (Y < H)[:, :, None]
print("instances", I, "pixel-pairs now", I * 256, "alive pixel pairs", int(alive.sum()), "=%.3f" % (alive.sum() / (I * 256)))
for (rh, rw) in ((16, 16), (16, 8), (8, 8), (4, 8), (4, 4)):
    a = alive.reshape(I, 16 // rh, rh, 16 // rw, rw).any(axis=(2, 4))
    nrec = int(a.sum())
    print("region %2dx%-2d (h x w): records %8d (%.2fx of I)  pixel-pair evals %.3f of now" % (rh, rw, nrec, nrec / I, nrec * rh * rw / (I * 256)))
