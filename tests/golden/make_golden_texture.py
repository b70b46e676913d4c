"""Generate tests/golden/texture.npz by EXECUTING the reference's duplicate_texture_vertex_color_2 (helpers.py:930-941) and
process_uv (helpers.py:945-950), cut out with `ast`, on a synthetic seam topology (build container only).

    python tests/golden/make_golden_texture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, cut  # noqa: E402


def main():
    rng = np.random.default_rng(20261020)
    ns = {"np": np}
    exec(cut(os.path.join(REF, "helpers.py"), {"duplicate_texture_vertex_color_2", "process_uv"}), ns)
    n_vert = 200
    # every vertex owns 1..3 UV coordinates (seam vertices own several); uvs_ori lists them all, shuffled
    per = rng.integers(1, 4, n_vert)
    uvs_texture = [[tuple(np.round(rng.uniform(0, 1, 2), 6)) for _ in range(k)] for k in per]
    flat = [(uv, i) for i, l in enumerate(uvs_texture) for uv in l]
    order = rng.permutation(len(flat))
    uvs_ori = np.array([flat[j][0] for j in order])
    variables = {"uvs_ori": uvs_ori, "uvs_texture_ori": uvs_texture}
    colors = rng.uniform(0, 1, (n_vert, 3))
    ref = np.array(ns["duplicate_texture_vertex_color_2"](variables, colors))
    uv_in = rng.uniform(0, 1, (50, 2))
    ref_uv = ns["process_uv"](uv_in.copy(), 1024, 1024)
    np.savez_compressed(os.path.join(HERE, "texture.npz"), uvs_ori=uvs_ori, per=per,
                        uvs_texture_flat=np.array([uv for l in uvs_texture for uv in l]), colors=colors, ref_colors=ref,
                        uv_in=uv_in, ref_uv=ref_uv)
    print("texture.npz written", ref.shape, ref_uv.shape)


if __name__ == "__main__":
    main()
