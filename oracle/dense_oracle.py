"""CPU oracle of the dense-mesh attribute interpolation -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates compute_vertex_attribute_by_weight_2 (reference helpers.py:237-253) with explicit loops over the four quad
corners (float64 products, left-to-right sum: the order NumPy's np.sum(axis=1) uses for this [M,4,C] array), followed by
the float32 cast its caller applies (train.py:504-506 `.cuda().float()`).  PARITY PINNED: tests check it bit for bit
against tests/golden/dense.npz, produced by executing the reference function itself (tests/golden/make_golden_dense.py).
"""
import numpy as np


def compute_vertex_attribute_by_weight_2(variables, attribute):
    father = np.asarray(variables["dense_vertex_father"]).reshape(-1)
    weight = np.asarray(variables["dense_vertex_weight"], dtype=np.float64)
    quads = np.asarray(variables["dense_quad_faces"])
    attribute = np.asarray(attribute)
    n_dense = variables["dense_vertex"].shape[0]
    out = np.zeros((n_dense, attribute.shape[1]), np.float64)
    out[:attribute.shape[0]] = attribute
    corners = quads[father]                                  # [M,4]
    acc = attribute[corners[:, 0]].astype(np.float64) * weight[:, 0:1]
    for j in range(1, 4):
        acc = acc + attribute[corners[:, j]].astype(np.float64) * weight[:, j:j + 1]
    out[attribute.shape[0]:] = acc
    return out.astype(np.float32)
