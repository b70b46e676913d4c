"""CPU: the loss / Adam oracle (oracle/loss_oracle.py) against the golden vectors the REFERENCE'S OWN code produced
(tests/golden/make_golden_loss.py: calc_ssim/_ssim/create_window/gaussian of external.py:71-116, l1_loss_v1 of
helpers.py:115-116, the get_loss expression of train.py:310,317, torch.optim.Adam as built at train.py:272-297)."""
import os

import numpy as np

from oracle import loss_oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("a", "b", "c", "d")


def rel_err(a, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(np.asarray(a, np.float64) - ref) / (np.abs(ref) + 1e-3 * np.max(np.abs(ref)) + 1e-30)))


def test_window_matches_reference_gaussian():
    g = np.load(os.path.join(G, "loss.npz"))
    np.testing.assert_allclose(loss_oracle.gaussian_window().numpy(), g["window_1d"], rtol=2e-7, atol=0)
    assert abs(float(g["window_1d"].sum()) - 1.0) < 1e-6


def test_image_loss_oracle_matches_reference_outputs():
    g = np.load(os.path.join(G, "loss.npz"))
    for n in CASES:
        affine = bool(g[n + "_affine"])
        o = loss_oracle.image_loss(g[n + "_render"][None], g[n + "_target"][None],
                                   g[n + "_cam_m"][None] if affine else None, g[n + "_cam_c"][None] if affine else None)
        # the golden run is the reference in float32; the oracle is float64: values to 2e-6, gradients to 3e-4 relative
        np.testing.assert_allclose(o["terms"][0, :3], g[n + "_terms"], rtol=0, atol=2e-6)
        assert rel_err(o["d_render"][0], g[n + "_d_render"]) < 3e-4, n
        if affine:
            assert rel_err(o["d_cam_m"][0], g[n + "_d_cam_m"]) < 3e-4
            assert rel_err(o["d_cam_c"][0], g[n + "_d_cam_c"]) < 3e-4


def test_image_loss_oracle_properties():
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (2, 3, 20, 31))
    o = loss_oracle.image_loss(x, x)
    np.testing.assert_allclose(o["terms"][:, 0], 0, atol=1e-15)          # L1 of identical images
    np.testing.assert_allclose(o["terms"][:, 1], 1, atol=1e-12)          # SSIM of identical images
    assert np.abs(o["d_render"]).max() < 1e-9                            # sign(0) = 0 and SSIM is at its maximum
    # views are independent: the batch equals the per-view calls
    y = rng.uniform(0, 1, x.shape)
    both = loss_oracle.image_loss(x, y)
    one = loss_oracle.image_loss(x[1:], y[1:])
    np.testing.assert_allclose(both["terms"][1], one["terms"][0], rtol=1e-12)
    np.testing.assert_allclose(both["d_render"][1], one["d_render"][0], rtol=1e-12, atol=1e-18)


def test_adam_oracle_matches_torch_optim_adam_as_the_reference_builds_it():
    g = np.load(os.path.join(G, "adam.npz"))
    lrs = {"means3D": 0.0, "rgb_colors": 0.0025, "unnorm_rotations": 0.001, "log_scales": 0.001, "cam_m": 1e-4}
    steps, change = int(g["steps"]), int(g["lr_change_step"])
    for k, lr in lrs.items():
        grads = g[k + "_grads"]
        if k == "means3D":           # lr 0 for the first `change` steps, then 1.6e-5 (update_optimizer): moments run on
            p, m, v = np.array(g[k + "_init"], np.float64), 0.0, 0.0
            for s in range(steps):
                m = 0.9 * m + 0.1 * grads[s].astype(np.float64)
                v = 0.999 * v + 0.001 * grads[s].astype(np.float64) ** 2
                cur = 0.0 if s < change else 0.000016
                p = p - (cur / (1 - 0.9 ** (s + 1))) * m / (np.sqrt(v) / np.sqrt(1 - 0.999 ** (s + 1)) + 1e-15)
        else:
            p, m, v = loss_oracle.adam_steps(g[k + "_init"], grads, lr)
        np.testing.assert_allclose(p, g[k + "_final"], rtol=2e-6, atol=2e-7)
        assert rel_err(m, g[k + "_exp_avg"]) < 5e-5 and rel_err(v, g[k + "_exp_avg_sq"]) < 5e-5      # golden is fp32


def test_dense_attribute_oracle_is_bit_identical_to_the_reference_function():
    from oracle import dense_oracle
    g = np.load(os.path.join(G, "dense.npz"))
    for n in ("small", "wide"):
        var = {k: g[f"{n}_{k}"] for k in ("dense_quad_faces", "dense_vertex_father", "dense_vertex_weight", "dense_vertex")}
        out = dense_oracle.compute_vertex_attribute_by_weight_2(var, g[n + "_attr"])
        np.testing.assert_array_equal(out, g[n + "_ref_cuda_float"])
        np.testing.assert_array_equal(out[:g[n + "_attr"].shape[0]], g[n + "_attr"])


def test_activations_golden_is_reproduced_by_the_reference_expression():
    """helpers.py:91-100 restated as plain PyTorch reproduces the fixture the reference function itself generated (this
    also exercises the shared checker the GPU test applies to the fused kernel)."""
    import torch
    from tests import activations_check

    def params2rendervar(params):
        return {"means3D": params["means3D"], "colors_precomp": params["rgb_colors"],
                "rotations": torch.nn.functional.normalize(params["unnorm_rotations"]),
                "opacities": torch.sigmoid(params["logit_opacities"]), "scales": torch.exp(params["log_scales"]),
                "means2D": torch.zeros_like(params["means3D"], requires_grad=True) + 0}
    activations_check.run_and_check(params2rendervar, "cpu")
    # a zero LEAF for means2D (what the fused mirror returns) satisfies the same contract
    def with_leaf(params):
        rv = params2rendervar(params)
        rv["means2D"] = torch.zeros_like(params["means3D"], requires_grad=True)
        return rv
    activations_check.run_and_check(with_leaf, "cpu")
