"""Whole-iteration CUDA graphs ("CUDA-graph the whole step").

In the reference's regime (one 512x375 view and ~8k Gaussians per Adam step, train.py:661-673) an iteration is a few
hundred microseconds of kernels behind ~100 tiny launches, so the host, not the GPU, sets the pace.  Everything on this
repository's path is capture-safe -- the rasterizer never touches the host while a stream is capturing, the fused image
loss has no host round trip, and ``FusedAdam(capturable=True)`` keeps its step counters and learning rates on the
device -- so a whole iteration (render -> loss -> backward -> optimiser step) can be recorded once and replayed:

    opt = FusedAdam(groups, lr=0.0, eps=1e-15, capturable=True)
    def iteration():
        im, radius, _, _ = Renderer(raster_settings=cam)(**params2rendervar(params))
        loss = image_loss(im, gt, params['cam_m'][i], params['cam_c'][i])
        loss.backward()
        opt.step(); opt.zero_grad(set_to_none=True)
        return loss
    step = capture(iteration)            # a few eager warm-up runs, then one capture
    for _ in range(n): step.replay()     # step.outputs is the (static) loss tensor

Static-shape rules of CUDA graphs apply: tensors the iteration reads (ground-truth image, camera block, parameters) must
keep their storage -- update them in place (``gt.copy_(next_frame)``); one graph per camera is the natural granularity.
The instance capacity of every captured render is fixed at capture time; ``check()`` reads the device status blocks
(synchronising) and raises if a replay overflowed it, after raising the remembered capacity for a re-capture.
"""
from __future__ import annotations

import torch

from . import engine, rasterizer


class CapturedStep:
    def __init__(self, graph: torch.cuda.CUDAGraph, outputs, states):
        self.graph, self.outputs, self.states = graph, outputs, states
        self.replays = 0

    def replay(self):
        self.graph.replay()
        self.replays += 1
        return self.outputs

    def check(self) -> None:
        """Synchronising validation of every render inside the graph (call it occasionally, e.g. once per frame)."""
        for st in self.states:
            st._status = None
            s = st.status()
            if s.overflow:
                engine._CAP_MEMO[st.key] = int(s.num_instances * 1.5) + 4096
                raise RuntimeError("topo4d_b200: a replayed render needed %d (tile, Gaussian) instances but the captured "
                                   "workspace holds %d; re-capture the step (the capacity has been raised)"
                                   % (s.num_instances, s.cap_instances))


def capture(fn, warmup: int = 3, pool=None, capacity_headroom: float = 2.0) -> CapturedStep:
    """Run `fn()` `warmup` times eagerly on a side stream (sizes workspaces, creates optimiser state), then record one
    more call into a CUDA graph.  `fn` must be free of host synchronisation and read only statically allocated tensors.
    `capacity_headroom`: the captured renders get room for this many times the (tile, Gaussian) instances the warm-up
    needed -- the capacity cannot grow during replays, and optimisation moves / resizes the splats."""
    dev = torch.cuda.current_device()
    engine._TOUCHED_KEYS.clear()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(max(warmup, 1)):
            fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    rasterizer._check_pending()
    for key in engine._TOUCHED_KEYS:
        if key in engine._COUNT_MEMO:          # room relative to the MEASURED count (never compounds across captures)
            engine._CAP_MEMO[key] = max(engine._CAP_MEMO.get(key, 0), int(engine._COUNT_MEMO[key] * capacity_headroom) + 4096)
    del rasterizer._CAPTURE_LOG[:]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, pool=pool):
        out = fn()
    states = list(rasterizer._CAPTURE_LOG)
    del rasterizer._CAPTURE_LOG[:]
    return CapturedStep(g, out, states)
