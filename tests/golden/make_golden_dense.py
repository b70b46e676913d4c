"""Generate tests/golden/dense.npz by EXECUTING the reference's compute_vertex_attribute_by_weight_2
(helpers.py:237-253, cut out with `ast`) on a synthetic densified quad mesh (build container only).

    python tests/golden/make_golden_dense.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, cut  # noqa: E402


def synthetic_topology(rng, n_base, n_quads, per_quad):
    quads = np.stack([rng.permutation(n_base)[:4] for _ in range(n_quads)]).astype(np.int64)
    m = n_quads * per_quad
    father = np.repeat(np.arange(n_quads, dtype=np.int32), per_quad)[:, None]          # [M,1] int32 (helpers.py:612)
    u, v = rng.uniform(0, 1, m), rng.uniform(0, 1, m)
    weight = np.stack([(1 - u) * (1 - v), u * (1 - v), u * v, (1 - u) * v], 1)           # bilinear, float64 (helpers.py:613)
    return {"dense_quad_faces": quads, "dense_vertex_father": father, "dense_vertex_weight": weight,
            "dense_vertex": np.zeros((n_base + m, 3))}


def main():
    rng = np.random.default_rng(20261019)
    ns = {"np": np, "torch": torch}
    exec(cut(os.path.join(REF, "helpers.py"), {"compute_vertex_attribute_by_weight_2"}), ns)
    out = {}
    for name, (n_base, n_quads, per_quad, ch) in {"small": (50, 12, 9, 3), "wide": (300, 200, 16, 5)}.items():
        var = synthetic_topology(rng, n_base, n_quads, per_quad)
        attr = rng.normal(0, 1, (n_base, ch)).astype(np.float32)                          # .cpu().numpy() of a float32 parameter
        ref = ns["compute_vertex_attribute_by_weight_2"](var, attr)                      # float64 [n_dense, ch]
        out[name + "_attr"] = attr
        for k, v in var.items():
            out[name + "_" + k] = v
        out[name + "_ref_f64"] = ref
        out[name + "_ref_cuda_float"] = torch.from_numpy(ref).float().numpy()            # train.py:504-506
    np.savez_compressed(os.path.join(HERE, "dense.npz"), **out)
    print("dense.npz written")


if __name__ == "__main__":
    main()
