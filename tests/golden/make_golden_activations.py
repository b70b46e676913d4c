"""Generate tests/golden/activations.npz by EXECUTING the reference's params2rendervar (helpers.py:91-100, cut out with
`ast`; only device="cuda" is neutralised) on the CPU, with torch.autograd gradients under fixed upstream weights.

    python tests/golden/make_golden_activations.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, cut  # noqa: E402


def main():
    rng = np.random.default_rng(20261020)
    ns = {"torch": torch}
    exec(cut(os.path.join(REF, "helpers.py"), {"params2rendervar"}).replace(', device="cuda"', ""), ns)
    n = 257
    raw = {"means3D": rng.normal(0, 0.3, (n, 3)), "rgb_colors": rng.uniform(0, 1, (n, 3)),
           "unnorm_rotations": rng.normal(0, 2.0, (n, 4)), "logit_opacities": rng.normal(0, 4.0, (n, 1)),
           "log_scales": rng.normal(-3.5, 1.0, (n, 3))}
    raw["unnorm_rotations"][5] = 0.0                       # the eps-clamped branch of F.normalize
    raw["logit_opacities"][6] = 1000.0                     # sigmoid(1000) = 1 (train.py:142)
    raw = {k: v.astype(np.float32) for k, v in raw.items()}
    params = {k: torch.tensor(v, requires_grad=True) for k, v in raw.items()}
    rv = ns["params2rendervar"](params)
    w = {k: rng.normal(0, 1, tuple(rv[k].shape)).astype(np.float32) for k in ("rotations", "opacities", "scales")}
    sum((rv[k] * torch.tensor(w[k])).sum() for k in w).backward()
    out = {"in_" + k: v for k, v in raw.items()}
    out.update({"w_" + k: v for k, v in w.items()})
    out.update({"out_" + k: rv[k].detach().numpy() for k in ("rotations", "opacities", "scales", "means2D")})
    out.update({"grad_" + k: params[k].grad.numpy() for k in ("unnorm_rotations", "logit_opacities", "log_scales")})
    out["keys"] = np.array(sorted(rv.keys()))
    np.savez_compressed(os.path.join(HERE, "activations.npz"), **out)
    print("activations.npz written; keys:", sorted(rv.keys()))


if __name__ == "__main__":
    main()
