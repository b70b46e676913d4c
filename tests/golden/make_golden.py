"""Generate the committed golden vectors under tests/golden/ FROM THE REFERENCE'S OWN PYTHON.

Run in the build container (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

The reference modules cannot be imported whole (helpers.py / external.py pull in open3d,
pywavefront, trimesh, skimage, nvdiffrast and the un-vendored CUDA rasterizer), so the specific
reference functions are cut out of their source files with `ast` and executed unmodified, except
that `.cuda()` / device='cuda' are neutralised (no GPU here).  Outputs:

  sh_eval.npz        eval_sh (helpers.py:867-922) + constants (helpers.py:836-864)
  rotation.npz       build_rotation (external.py:26-43), build_quaterion (external.py:45-61)
  camera.npz         setup_camera (helpers.py:63-88): viewmatrix / projmatrix / tanfov / campos
  face3d_small.npz   reference face3d _render_colors_core (mesh_core.cpp:169-234) outputs on small
                     meshes, via oracle/_ref/libf3d_ref.so (built by oracle/Makefile from the
                     reference sources where they lie)
"""
import ast
import collections
import os
import sys

import numpy as np
import torch

REF = os.environ.get("TOPO4D_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def cut(path, names):
    """Source text of the top-level defs/assignments called `names` in `path`."""
    src = open(path).read()
    tree = ast.parse(src)
    out = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            out.append(ast.get_source_segment(src, node))
        elif isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id in names for t in node.targets):
            out.append(ast.get_source_segment(src, node))
    return "\n\n".join(out)


def main():
    rng = np.random.default_rng(20261017)

    # ---- SH ----
    ns = {"np": np, "torch": torch}
    exec(cut(os.path.join(REF, "helpers.py"), {"C0", "C1", "C2", "C3", "C4", "eval_sh", "RGB2SH", "SH2RGB"}), ns)
    n = 64
    sh = rng.normal(0, 0.5, (n, 16, 3))
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    out = {"sh": sh, "dirs": dirs, "C0": ns["C0"], "C1": ns["C1"], "C2": np.array(ns["C2"]), "C3": np.array(ns["C3"])}
    for deg in range(4):
        # eval_sh takes [..., C, K]; the op takes [N, K, 3]
        out[f"eval_deg{deg}"] = ns["eval_sh"](deg, np.transpose(sh, (0, 2, 1)), dirs)
    out["rgb2sh_of_half"] = ns["RGB2SH"](np.array([0.25, 0.5, 0.75]))
    np.savez(os.path.join(HERE, "sh_eval.npz"), **out)

    # ---- rotations ----
    ns = {"torch": torch}
    src = cut(os.path.join(REF, "external.py"), {"build_rotation", "build_quaterion"}).replace(", device='cuda'", "")
    exec(src, ns)
    q = rng.normal(size=(32, 4))
    R = ns["build_rotation"](torch.tensor(q, dtype=torch.float64)).numpy()
    nrm = rng.normal(size=(32, 3))
    bq = ns["build_quaterion"](torch.tensor(nrm, dtype=torch.float32)).numpy()
    np.savez(os.path.join(HERE, "rotation.npz"), q=q, R=R, normals=nrm, build_quaterion=bq)

    # ---- camera ----
    Camera = collections.namedtuple("Camera", "image_height image_width tanfovx tanfovy bg scale_modifier viewmatrix "
                                              "projmatrix sh_degree campos prefiltered debug")
    ns = {"torch": torch, "Camera": Camera}
    src = cut(os.path.join(REF, "helpers.py"), {"setup_camera"}).replace(".cuda()", "").replace(', device="cuda"', "")
    exec(src, ns)
    from topo4d_b200 import synth
    cams = {}
    for i, (eye, w, h, fx, fy, cx, cy) in enumerate([((0.3, -0.2, -2.0), 512, 375, 900.0, 880.0, 250.0, 190.0),
                                                    ((1.0, 0.5, 1.5), 1920, 1080, 1728.0, 1728.0, 960.0, 540.0)]):
        w2c = synth.look_at(eye).astype(np.float32)
        k = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
        cam = ns["setup_camera"](None, w, h, k, w2c, near=0.01, far=100)
        pts = rng.uniform(-0.3, 0.3, (16, 3))
        pc = (w2c[:3, :3].astype(np.float64) @ pts.T).T + w2c[:3, 3]
        pix = np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], 1)   # pinhole ground truth
        cams[f"c{i}_w2c"] = w2c
        cams[f"c{i}_whk"] = np.array([w, h, fx, fy, cx, cy], np.float64)
        cams[f"c{i}_viewmatrix"] = cam.viewmatrix.contiguous().numpy().reshape(4, 4)
        cams[f"c{i}_projmatrix"] = cam.projmatrix.contiguous().numpy().reshape(4, 4)
        cams[f"c{i}_tanfov"] = np.array([cam.tanfovx, cam.tanfovy], np.float64)
        cams[f"c{i}_campos_ref"] = cam.campos.numpy()            # the reference's (always-zero) campos, helpers.py:66
        cams[f"c{i}_points"] = pts
        cams[f"c{i}_pinhole_pix"] = pix
    np.savez(os.path.join(HERE, "camera.npz"), **cams)

    # ---- face3d (reference C++ compiled from where it lies) ----
    from oracle import f3d_oracle
    if not f3d_oracle.have_ref():
        raise SystemExit("oracle/_ref/libf3d_ref.so missing: run `make -C oracle ref` in the build container")
    f3d = {}
    cases = {
        "tri_corner": (np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0]], np.float64), np.array([[0, 1, 2]]), 24, 20),
        "coincident": (np.array([[3, 3, 1], [15, 4, 1], [5, 14, 1]], np.float64), np.array([[0, 1, 2], [0, 1, 2], [2, 1, 0]]), 20, 20),
    }
    v, t, c = synth.uv_grid_mesh(grid=9, res=64, seed=3)
    cases["grid9_64"] = (v, t, 64, 64)
    v2, t2, _ = synth.uv_grid_mesh(grid=5, res=40, seed=4)
    v2[:, 2] = rng.normal(size=v2.shape[0])                       # real depth test
    cases["grid5_depth"] = (v2, t2, 40, 40)
    for name, (vv, tt, h, w) in cases.items():
        col = rng.uniform(0, 1, (vv.shape[0], 3))
        img, dep = f3d_oracle.render_colors_ref(vv, tt, col, h, w, 3)
        f3d[name + "_vertices"] = vv
        f3d[name + "_triangles"] = tt
        f3d[name + "_colors"] = col
        f3d[name + "_hw"] = np.array([h, w])
        f3d[name + "_image"] = img
        f3d[name + "_depth"] = dep
    np.savez_compressed(os.path.join(HERE, "face3d_small.npz"), **f3d)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
