"""Host-side mirror of the reference's ``params2rendervar`` (helpers.py:91-112) over the C ABI.

The reference applies the parameter activations outside the rasterizer, once per iteration:
    'rotations': torch.nn.functional.normalize(params['unnorm_rotations'])
    'opacities': torch.sigmoid(params['logit_opacities'])
    'scales':    torch.exp(params['log_scales'])
    'means2D':   torch.zeros_like(params['means3D'], requires_grad=True, device="cuda") + 0
which costs PyTorch ~7 launches forward and ~12 backward on tiny tensors.  ``params2rendervar`` / ``params2rendervar_dense``
below return the same dictionaries (same keys, shapes, gradients) from ONE forward and ONE backward launch
(csrc/t4d_activate.cu).  ``means2D`` is returned as a zero LEAF that requires grad (``retain_grad()`` on it is a no-op and
``.grad`` is populated by the rasterizer's backward, which is all the reference uses it for, train.py:304,311).  CUDA-only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _c(t):
    return t if (t.dtype is torch.float32 and t.is_contiguous()) else t.float().contiguous()


class _Activate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, unnorm_rotations, logit_opacities, log_scales):
        if not unnorm_rotations.is_cuda:
            raise RuntimeError("topo4d_b200: params2rendervar is CUDA-only; there is no CPU path")
        dev = unnorm_rotations.device
        q, l, s = _c(unnorm_rotations), _c(logit_opacities), _c(log_scales)
        n = int(q.shape[0])
        if q.shape != (n, 4) or l.numel() != n or s.shape != (n, 3):
            raise ValueError("expected unnorm_rotations [N,4], logit_opacities [N,1], log_scales [N,3]")
        rot, opac, sc = torch.empty_like(q), torch.empty_like(l), torch.empty_like(s)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().t4d_activate(_p(q), _p(l), _p(s), n, _p(rot), _p(opac), _p(sc),
                                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "t4d_activate")
        ctx.save_for_backward(q, opac, sc)
        return rot, opac, sc

    @staticmethod
    def backward(ctx, g_rot, g_opac, g_sc):
        q, opac, sc = ctx.saved_tensors
        dev, n = q.device, int(q.shape[0])
        need = ctx.needs_input_grad
        g_rot = None if g_rot is None else _c(g_rot)
        g_opac = None if g_opac is None else _c(g_opac)
        g_sc = None if g_sc is None else _c(g_sc)
        d_q = torch.empty_like(q) if need[0] else None
        d_l = torch.empty_like(opac) if need[1] else None
        d_s = torch.empty_like(sc) if need[2] else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().t4d_activate_backward(_p(q), _p(opac), _p(sc), _p(g_rot), _p(g_opac), _p(g_sc), n,
                                                        _p(d_q), _p(d_l), _p(d_s),
                                                        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "t4d_activate_backward")
        return d_q, d_l, d_s


def activate(unnorm_rotations, logit_opacities, log_scales):
    """(normalize(unnorm_rotations), sigmoid(logit_opacities), exp(log_scales)) with autograd, one launch each way."""
    return _Activate.apply(unnorm_rotations, logit_opacities, log_scales)


def params2rendervar(params):
    """helpers.py:91-100 -- same dictionary, fused activations."""
    rot, opac, sc = activate(params["unnorm_rotations"], params["logit_opacities"], params["log_scales"])
    return {"means3D": params["means3D"], "colors_precomp": params["rgb_colors"], "rotations": rot, "opacities": opac,
            "scales": sc, "means2D": torch.zeros_like(params["means3D"], requires_grad=True)}


def params2rendervar_dense(params, variables=None):
    """helpers.py:102-112 -- the dense Gaussian mesh's dictionary."""
    rot, opac, sc = activate(params["dense_unnorm_rotations"], params["dense_logit_opacities"], params["dense_log_scales"])
    return {"means3D": params["dense_means3D"], "colors_precomp": params["dense_rgb_colors"], "rotations": rot,
            "opacities": opac, "scales": sc, "means2D": torch.zeros_like(params["dense_means3D"], requires_grad=True)}
