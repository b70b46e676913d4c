import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diff_gaussian_rasterization import GaussianRasterizationSettings as Camera, GaussianRasterizer as Renderer
from topo4d_b200 import synth, engine
n, w, h = 8280, 512, 375
sc = synth.head_scene(n, seed=0, sh_degree=None, opacity="topo4d")
c = synth.ring_cameras(24, w=w, h=h, radius=0.6, focal_over_h=1.6)[0]
dev = "cuda"
params = {k: torch.tensor(v, device=dev, requires_grad=True) for k, v in sc.items()}
w2c = torch.tensor(c.w2c, dtype=torch.float32, device=dev)
cam = Camera(image_height=h, image_width=w, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
             viewmatrix=w2c.unsqueeze(0).transpose(1, 2), projmatrix=torch.tensor(c.projmatrix, device=dev).unsqueeze(0), sh_degree=0,
             campos=torch.tensor(c.campos, device=dev), prefiltered=False, debug=False)
target = torch.rand(3, h, w, device=dev)
T = {"rendervar": 0.0, "render": 0.0, "loss": 0.0, "backward": 0.0}
def it(acc):
    t0 = time.perf_counter()
    rv = {"means3D": params["means3D"], "colors_precomp": params["colors_precomp"], "rotations": torch.nn.functional.normalize(params["rotations"]),
          "opacities": params["opacities"], "scales": params["scales"], "means2D": torch.zeros_like(params["means3D"], requires_grad=True) + 0}
    t1 = time.perf_counter()
    im, radius, _, _ = Renderer(raster_settings=cam)(**rv)
    t2 = time.perf_counter()
    loss = (im - target).abs().mean()
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    for p in params.values(): p.grad = None
    if acc:
        T["rendervar"] += t1 - t0; T["render"] += t2 - t1; T["loss"] += t3 - t2; T["backward"] += t4 - t3
for i in range(50): it(False)
torch.cuda.synchronize()
N = 300
for i in range(N): it(True)
torch.cuda.synchronize()
print("CPU us per iter:", {k: round(v / N * 1e6, 1) for k, v in T.items()}, "SYNC=", os.environ.get("TOPO4D_B200_SYNC", "1"))
# kernel-only time of our op for this workload
t = {k: v.detach() for k, v in params.items()}
camt = torch.tensor(engine.pack_cameras_numpy([c]), device=dev)
ev = {}
for i in range(20):
    out = engine.forward(t["means3D"], t["opacities"], camt, h, w, colors_precomp=t["colors_precomp"], scales=t["scales"], rotations=t["rotations"], check="none", stage_events=ev)
    engine.backward(out[4], torch.ones_like(out[0]), stage_events=ev)
torch.cuda.synchronize()
print("our kernels us per launch:", {k: round(float(np.mean([a.elapsed_time(b) for a, b in v[5:]])) * 1e3, 1) for k, v in ev.items()})
