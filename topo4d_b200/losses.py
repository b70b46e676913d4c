"""Host-side mirror of the reference's per-iteration image loss over the C ABI.

Reference code this replaces (same names, argument meaning and reductions):
  * ``l1_loss_v1(x, y)``                       helpers.py:115-116   mean |x - y|
  * ``calc_ssim(img1, img2, window_size=11)``  external.py:85-116   11x11 Gaussian-window SSIM, zero padding, mean
  * the expression in ``get_loss``             train.py:310,317
        im = exp(cam_m[id])[:, None, None] * im + cam_c[id][:, None, None]
        loss = 0.8 * l1_loss_v1(im, gt) + 0.2 * (1.0 - calc_ssim(im, gt))
``image_loss`` evaluates that whole expression AND its backward in one call of ``t4d_image_loss`` (two tile kernels
plus a finalize, csrc/t4d_loss.cu); PyTorch is used for memory and autograd plumbing only.  CUDA-only: CPU tensors
raise, there is no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_WS_BYTES: dict[tuple, int] = {}


def _f32c(t, dev):
    if t is None:
        return None
    if t.dtype is torch.float32 and t.device == dev and t.is_contiguous():
        return t
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _run(render, target, cam_m, cam_c, w_l1, w_ssim, need_grad):
    """render/target [V,3,H,W] fp32 CUDA contiguous -> (loss[V,4], dL_drender | None, dL_dcam_m | None, dL_dcam_c | None)."""
    if not render.is_cuda:
        raise RuntimeError("topo4d_b200: the image loss is CUDA-only (input is on %s); there is no CPU path" % render.device)
    dev = render.device
    V, Cn, H, W = render.shape
    if Cn != 3 or tuple(target.shape) != (V, 3, H, W):
        raise ValueError(f"image loss expects [V,3,H,W] render and target of equal shape, got {tuple(render.shape)} and {tuple(target.shape)}")
    L = _lib.lib()
    key = (V, H, W)
    nbytes = _WS_BYTES.get(key)
    if nbytes is None:
        nbytes = _WS_BYTES[key] = L.t4d_image_loss_workspace_bytes(V, H, W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    loss = torch.empty((V, 4), dtype=torch.float32, device=dev)
    d_render = torch.empty_like(render) if need_grad else None
    has_cam = cam_m is not None
    d_m = torch.empty((V, 3), dtype=torch.float32, device=dev) if need_grad and has_cam else None
    d_c = torch.empty((V, 3), dtype=torch.float32, device=dev) if need_grad and has_cam else None
    args = _lib.T4dImageLoss(V, H, W, float(w_l1), float(w_ssim), _ptr(render), _ptr(target), _ptr(cam_m), _ptr(cam_c),
                             _ptr(loss), _ptr(d_render), _ptr(d_m), _ptr(d_c), _ptr(ws), nbytes)
    with torch.cuda.device(dev):
        _lib.check(L.t4d_image_loss(C.byref(args), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "t4d_image_loss")
    return loss, d_render, d_m, d_c


class _ImageLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, target, cam_m, cam_c, w_l1, w_ssim):
        dev = render.device
        squeeze = render.dim() == 3
        r4 = _f32c(render.unsqueeze(0) if squeeze else render, dev)
        t4 = _f32c(target.unsqueeze(0) if target.dim() == 3 else target, dev)
        V = r4.shape[0]
        m2 = None if cam_m is None else _f32c(cam_m.reshape(V, 3), dev)
        c2 = None if cam_c is None else _f32c(cam_c.reshape(V, 3), dev)
        need = render.requires_grad or (cam_m is not None and cam_m.requires_grad) or (cam_c is not None and cam_c.requires_grad)
        loss, d_r, d_m, d_c = _run(r4, t4, m2, c2, w_l1, w_ssim, need)
        ctx.shapes = (render.shape, None if cam_m is None else cam_m.shape, None if cam_c is None else cam_c.shape)
        ctx.save_for_backward(d_r, d_m, d_c)
        ctx.mark_non_differentiable(loss)
        return loss[:, 2].sum(), loss

    @staticmethod
    def backward(ctx, g, _g_terms):
        d_r, d_m, d_c = ctx.saved_tensors
        rs, ms, cs = ctx.shapes
        if d_r is None:
            return None, None, None, None, None, None
        # dL/dim does not depend on the loss value; an upstream factor other than 1 costs one extra pass over the image
        # (fold outer weights into w_l1 / w_ssim to avoid it)
        return (d_r.view(rs) * g, None, None if ms is None else (d_m * g).view(ms), None if cs is None else (d_c * g).view(cs),
                None, None)


def image_loss(render, target, cam_m=None, cam_c=None, w_l1=0.8, w_ssim=0.2, return_terms=False):
    """``w_l1 * l1_loss_v1(im, target) + w_ssim * (1 - calc_ssim(im, target))`` with ``im = exp(cam_m) * render + cam_c``
    per channel (train.py:310,317), summed over views when ``render`` is [V,3,H,W].  Differentiable w.r.t. render, cam_m,
    cam_c.  ``cam_m`` / ``cam_c``: [3] (or [V,3]) rows ALREADY selected for the view(s), i.e. ``params['cam_m'][curr_id]``.
    ``return_terms=True`` also returns the per-view [V,4] tensor (l1 mean, ssim mean, total, 0) for logging."""
    if (cam_m is None) != (cam_c is None):
        raise ValueError("provide both cam_m and cam_c or neither")
    total, terms = _ImageLossFn.apply(render, target, cam_m, cam_c, float(w_l1), float(w_ssim))
    return (total, terms) if return_terms else total


def l1_loss_v1(x, y):
    """helpers.py:115-116 -- ``torch.abs(x - y).mean()`` for [3,H,W] / [V,3,H,W] CUDA images (mean over ALL elements)."""
    v = 1 if x.dim() == 3 else x.shape[0]
    return image_loss(x, y, None, None, 1.0, 0.0) / v


def calc_ssim(img1, img2, window_size=11, size_average=True):
    """external.py:85-116 -- mean SSIM of [3,H,W] / [B,3,H,W] CUDA images (Gaussian window 11, sigma 1.5, zero padding)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("the fused kernel implements the reference's only call pattern: window_size=11, size_average=True")
    v = 1 if img1.dim() == 3 else img1.shape[0]
    # total = w_ssim * (1 - ssim) with w_ssim = -1  ->  ssim - 1, per view
    return image_loss(img1, img2, None, None, 0.0, -1.0) / v + 1.0
