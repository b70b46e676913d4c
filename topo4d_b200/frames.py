"""Frame data path: what ``get_dataset`` does per video frame (train.py:73-103, called at
train.py:653 for the geometry stage and train.py:722 at full resolution for the texture stage) without the serial host work.

Reference, per camera and frame: ``np.array(Image.open(f)) / 255.0`` (PIL decode on one core, float64), ``rotate_image`` =
``skimage.transform.rotate(im, k*90, resize=True)`` (camera.py:203-205; a bilinear warp evaluated at every pixel for what is a
pixel permutation), ``torch.tensor(im).float().cuda().permute(2, 0, 1)`` -- 24 times per frame, 12 Mpixel each in the texture
stage, all before the first iteration of the frame can start.

Here: the compressed bytes are read on a worker thread, decoded by nvJPEG on the GPU (``torchvision.io.decode_jpeg(device=...)``,
batched over the cameras), scaled to float32 / 255, rotated by k * 90 degrees as an index transform (``torch.rot90``: exact) and
returned as [3, H, W] CUDA tensors; :class:`FramePrefetcher` does this for frame t+1 on a side stream while frame t is being
optimised, so the data path leaves the critical path entirely.  PNG (the reference globs both, train.py:76) and anything
nvJPEG rejects fall back to a CPU decode of that file only.

Parity notes.  (1) nvJPEG and libjpeg (PIL) are different IDCT / chroma-upsampling implementations: decoded bytes may differ by
a few LSB (tests/test_frames.py bounds it at max 6/255, mean < 1/255 on synthetic photo-like images at quality 92; measured
5/255 and 0.74/255); PNG is lossless and bit-identical.  (2) ``skimage`` is not installed here, so its rotate could not be executed: for
angle = k*90 and resize=True it resamples the image on a grid that coincides with the source pixel centres, i.e. it equals
``np.rot90(im, k)`` (counter-clockwise) up to the warp's floating-point error; border pixels, where skimage blends with
cval = 0 if the grid lands a rounding error outside the image, are the place to diff if a user has skimage at hand.
"""
from __future__ import annotations

import glob
import os
import threading
from typing import Callable, Sequence

import numpy as np
import torch


def list_frame_files(data_dir: str, seq: str, frame: int, blacklist: Sequence[str] = ()) -> list[str]:
    """Same listing and order as get_dataset (train.py:76-77): sorted *.jpg, then sorted *.png, minus blacklisted cameras."""
    d = os.path.join(data_dir, seq, "%06d" % frame)
    names = sorted(glob.glob(os.path.join(d, "*.jpg"))) + sorted(glob.glob(os.path.join(d, "*.png")))
    return [f for f in names if not any(os.path.basename(f).startswith(b) for b in blacklist)]


def rotate90(im: torch.Tensor, k: int) -> torch.Tensor:
    """[C,H,W] image rotated counter-clockwise by k*90 degrees: what ``rotate(im, k*90, resize=True)`` (camera.py:203-205)
    computes, as an index transform (no resampling)."""
    k %= 4
    return im if k == 0 else torch.rot90(im, k, dims=(1, 2)).contiguous()


def _cpu_decode(path: str) -> torch.Tensor:
    from PIL import Image
    with Image.open(path) as im:
        a = np.array(im)
    if a.ndim == 2:
        a = a[:, :, None]
    return torch.from_numpy(np.ascontiguousarray(a)).permute(2, 0, 1)          # uint8 [C,H,W]


def decode_images(paths: Sequence[str], device, raw: Sequence[bytes] | None = None) -> list[torch.Tensor]:
    """Compressed files -> uint8 [C,H,W] CUDA tensors.  JPEGs go through nvJPEG in one batched call; the rest via PIL."""
    from torchvision.io import ImageReadMode, decode_jpeg
    dev = torch.device(device)
    if raw is None:
        raw = [open(p, "rb").read() for p in paths]
    out: list[torch.Tensor | None] = [None] * len(paths)
    jpg = [i for i, p in enumerate(paths) if p.lower().endswith((".jpg", ".jpeg"))]
    if jpg:
        try:
            bufs = [torch.frombuffer(bytearray(raw[i]), dtype=torch.uint8) for i in jpg]
            dec = decode_jpeg(bufs, device=dev, mode=ImageReadMode.UNCHANGED)
            for i, d in zip(jpg, dec):
                out[i] = d
        except Exception:  # noqa: BLE001  (CMYK / progressive corner cases nvJPEG refuses: decode those on the host)
            pass
    for i, p in enumerate(paths):
        if out[i] is None:
            out[i] = _cpu_decode(p).to(dev, non_blocking=True)
    return out  # type: ignore[return-value]


def load_frame(paths: Sequence[str], rot_k: Sequence[int], device="cuda", raw: Sequence[bytes] | None = None) -> list[torch.Tensor]:
    """One video frame: for every camera file the float32 [C,H,W] image in [0,1], rotated by rot_k[i]*90 degrees -- the `im`
    entries get_dataset builds (train.py:79-99)."""
    imgs = decode_images(paths, device, raw)
    return [rotate90(im.to(torch.float32).mul_(1.0 / 255.0), int(k)) for im, k in zip(imgs, rot_k)]


class FramePrefetcher:
    """Loads frame t+1 while frame t is optimised.

        pf = FramePrefetcher(lambda t: list_frame_files(root, seq, t + 1), rot_k, device)
        for t in range(frame_num):
            ims = pf.get(t)            # ready (or waits for the worker); starts loading t+1
            ... optimise frame t ...

    File reads and the decode launch happen on a worker thread, the device work on a side stream; ``get`` makes the
    current stream wait for that stream, so no host synchronisation is needed on the consumer side."""

    def __init__(self, files_of: Callable[[int], Sequence[str]], rot_k_of: Callable[[Sequence[str]], Sequence[int]] | Sequence[int],
                 device="cuda", num_frames: int | None = None):
        self.files_of, self.rot_k_of, self.device, self.num_frames = files_of, rot_k_of, torch.device(device), num_frames
        self.stream = torch.cuda.Stream(self.device)
        self._pending: dict[int, tuple[threading.Thread, dict]] = {}

    def _start(self, t: int) -> None:
        if t in self._pending or (self.num_frames is not None and t >= self.num_frames):
            return
        box: dict = {}

        def work():
            try:
                paths = list(self.files_of(t))
                rk = self.rot_k_of(paths) if callable(self.rot_k_of) else self.rot_k_of
                raw = [open(p, "rb").read() for p in paths]
                with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
                    box["ims"] = load_frame(paths, rk, self.device, raw)
                    box["event"] = torch.cuda.Event()
                    box["event"].record(self.stream)
                box["paths"] = paths
            except Exception as e:  # noqa: BLE001
                box["error"] = e

        th = threading.Thread(target=work, daemon=True)
        th.start()
        self._pending[t] = (th, box)

    def get(self, t: int) -> list[torch.Tensor]:
        self._start(t)
        th, box = self._pending.pop(t)
        th.join()
        if "error" in box:
            raise box["error"]
        torch.cuda.current_stream(self.device).wait_event(box["event"])
        for im in box["ims"]:
            im.record_stream(torch.cuda.current_stream(self.device))
        self._start(t + 1)
        return box["ims"]
