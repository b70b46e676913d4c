"""Fused Adam for Topo4D's parameter dictionary over the C ABI.

Mirrors how the reference builds and drives its optimiser:
  * ``torch.optim.Adam(param_groups, lr=0.0, eps=1e-15)`` with one group per named parameter, each with its own
    ``lr`` and a ``name`` key (train.py:272-297), so ``update_optimizer`` (helpers.py:801-804), which edits
    ``optimizer.param_groups[i]['lr']`` by name, works on this class unchanged;
  * ``optimizer.step()`` / ``optimizer.zero_grad(set_to_none=True)`` once per iteration (train.py:672-673);
  * the boolean-mask overwrites that follow every step (train.py:676-700) can be registered once with :meth:`pin`
    and are then applied inside the same kernel launch.
State layout is torch.optim.Adam's (``state[p]['step' | 'exp_avg' | 'exp_avg_sq']``), so ``state_dict`` round-trips.
All parameters of one ``step()`` go through ONE launch of ``t4d_adam_step`` (csrc/t4d_optim.cu).  CUDA-only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        # id(param) -> (weak reference to the parameter, mask, values): the weak reference proves at look-up time that the id
        # still belongs to the tensor the pin was registered for (an id can be reused after a tensor is freed)
        self._pins: dict[int, tuple] = {}
        # capturable=True: step counters and learning rates live on the device, so a step() captured in a CUDA graph
        # (topo4d_b200.graph.capture) advances correctly at every replay; after editing param_groups[i]['lr']
        # (update_optimizer, helpers.py:801-804) call sync_hyperparams() -- no re-capture needed.
        self.capturable = bool(capturable)
        self._lr_dev: torch.Tensor | None = None
        g0 = self.param_groups[0]
        if any(g["betas"] != g0["betas"] or g["eps"] != g0["eps"] for g in self.param_groups):
            raise ValueError("FusedAdam shares betas / eps across groups (as the reference does); only lr is per group")

    def pin(self, param: torch.Tensor, mask: torch.Tensor, values: torch.Tensor | None = None) -> None:
        """After every step, rows of `param` where `mask` (bool [rows]) is set are overwritten with the same rows of
        `values` (same shape as `param`; None = zeros) -- the fused form of ``params[name][mask] = const``
        (train.py:676-700).  Later calls for the same parameter replace the registration; ``mask=None`` removes it."""
        if mask is None:
            self._pins.pop(id(param), None)
            return
        rows = param.shape[0]
        if mask.numel() != rows:
            raise ValueError(f"pin(): mask has {mask.numel()} entries, the parameter has {rows} rows")
        m = mask.to(device=param.device).reshape(rows).to(torch.uint8).contiguous()
        v = None if values is None else values.to(device=param.device, dtype=torch.float32).expand_as(param).contiguous()
        import weakref
        self._pins[id(param)] = (weakref.ref(param), m, v)

    def pin_compact(self, param: torch.Tensor, overwrites) -> None:
        """The reference's own form (train.py:676-699): an ORDERED list of ``(mask, values)`` statements
        ``params[name][mask] = values`` where `values` is a scalar, a row, or the compact ``[mask.sum(), ...]`` tensor boolean
        indexing takes.  They are merged, in order (later statements win), into the one full-shape mask / value table pin() holds."""
        rows = param.shape[0]
        full_m = torch.zeros(rows, dtype=torch.bool, device=param.device)
        full_v = torch.zeros_like(param, dtype=torch.float32)
        for mask, values in overwrites:
            mb = mask.to(device=param.device)
            if mb.dtype is not torch.bool:                       # index arrays, like the reference's region masks
                mb = torch.zeros(rows, dtype=torch.bool, device=param.device).index_fill_(0, mb.long().reshape(-1), True)
            full_v[mb] = values if not torch.is_tensor(values) else values.to(device=param.device, dtype=torch.float32)
            full_m |= mb
        self.pin(param, full_m, full_v)

    def sync_hyperparams(self) -> None:
        """Capturable mode: push the groups' current learning rates to their device slots (one small H2D copy)."""
        if self._lr_dev is not None:
            host = torch.tensor([float(g["lr"]) for g in self.param_groups], dtype=torch.float32)
            self._lr_dev.copy_(host, non_blocking=False)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        segs, keep = [], []
        dev = None
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        if capturing and not self.capturable:
            raise RuntimeError("FusedAdam.step() inside CUDA-graph capture needs FusedAdam(..., capturable=True)")
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("topo4d_b200.FusedAdam is CUDA-only; there is no CPU path")
                if p.dtype is not torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous fp32 parameters")
                if dev is None:
                    dev = p.device
                elif p.device != dev:
                    raise RuntimeError("FusedAdam: all parameters must live on one device")
                grad = p.grad if (p.grad.is_contiguous() and p.grad.dtype is torch.float32) else p.grad.float().contiguous()
                st = self.state[p]
                if not st:
                    if capturing:
                        raise RuntimeError("run at least one eager step() before capturing (optimizer state is created lazily)")
                    st["step"] = torch.zeros((), dtype=torch.int32, device=p.device) if self.capturable else 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                step_dev = lr_dev = None
                if self.capturable:
                    if self._lr_dev is None or self._lr_dev.numel() != len(self.param_groups):      # add_param_group grows it
                        if capturing:
                            raise RuntimeError("run at least one eager step() before capturing")
                        self._lr_dev = torch.zeros(len(self.param_groups), dtype=torch.float32, device=p.device)
                    step_dev = st["step"].data_ptr()
                    lr_dev = self._lr_dev.data_ptr() + 4 * gi
                    step_host = 0
                else:
                    st["step"] = int(st["step"]) + 1
                    step_host = st["step"]
                pin = self._pins.get(id(p))
                if pin is not None:
                    if pin[0]() is not p:                        # stale entry of a freed tensor whose id was reused
                        del self._pins[id(p)]
                        pin = None
                    else:
                        pin = pin[1:]
                        if pin[0].numel() != p.shape[0] or (pin[1] is not None and pin[1].shape != p.shape):
                            raise RuntimeError("FusedAdam: a pinned parameter changed shape; register the pin again")
                rw = int(p.numel() // p.shape[0]) if (pin is not None and p.dim() > 0 and p.shape[0] > 0) else 1
                segs.append(_lib.T4dAdamSegment(p.data_ptr(), grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                                None if pin is None else pin[0].data_ptr(),
                                                None if pin is None or pin[1] is None else pin[1].data_ptr(),
                                                p.numel(), rw, step_host, float(g["lr"]), step_dev, lr_dev))
                keep.append(grad)
        if not segs:
            return loss
        if self.capturable and not capturing:
            self.sync_hyperparams()             # eager steps always see the current lrs
        b1, b2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        L = _lib.lib()
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for i in range(0, len(segs), _lib.T4D_ADAM_MAX_SEGMENTS):
                chunk = segs[i:i + _lib.T4D_ADAM_MAX_SEGMENTS]
                arr = (_lib.T4dAdamSegment * len(chunk))(*chunk)
                _lib.check(L.t4d_adam_step(arr, len(chunk), float(b1), float(b2), float(eps), stream), "t4d_adam_step")
        return loss
