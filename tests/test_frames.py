"""Frame data path (topo4d_b200/frames.py; reference get_dataset, train.py:73-103).  CPU part: file listing order and the
rotation index transform against NumPy; GPU part: nvJPEG decode against PIL within the stated tolerance, PNG bit-exact,
rotation, and the one-frame-ahead prefetcher."""
import os

import numpy as np
import pytest
import torch

from topo4d_b200 import frames


def _photo(h, w, seed):
    """Smooth, photo-like content (JPEG decoders disagree most on hard edges and noise, least on natural gradients)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    ch = []
    for c in range(3):
        f = sum(rng.uniform(20, 60) * np.sin(x / rng.uniform(15, 80) + rng.uniform(0, 6)) * np.cos(y / rng.uniform(15, 80) + rng.uniform(0, 6))
                for _ in range(4))
        ch.append(128 + f)
    return np.clip(np.stack(ch, -1), 0, 255).astype(np.uint8)


def _write_sequence(root, frames_n=3, cams=("A", "B", "C"), size=(96, 128)):
    from PIL import Image
    for t in range(1, frames_n + 1):
        d = os.path.join(root, "seq", "%06d" % t)
        os.makedirs(d, exist_ok=True)
        for i, c in enumerate(cams):
            Image.fromarray(_photo(size[0], size[1], 10 * t + i)).save(os.path.join(d, c + ".jpg"), quality=92)
        Image.fromarray(_photo(size[0], size[1], 99 + t)).save(os.path.join(d, "Z.png"))
        Image.fromarray(_photo(size[0], size[1], 7)).save(os.path.join(d, "skip_me.jpg"))


def test_listing_order_and_rotation_index_transform(tmp_path):
    _write_sequence(str(tmp_path), 1)
    files = frames.list_frame_files(str(tmp_path), "seq", 1, blacklist=["skip"])
    assert [os.path.basename(f) for f in files] == ["A.jpg", "B.jpg", "C.jpg", "Z.png"]      # sorted jpg, then sorted png (train.py:76)
    im = torch.arange(3 * 5 * 7, dtype=torch.float32).reshape(3, 5, 7)
    for k in range(-1, 5):
        got = frames.rotate90(im, k).numpy()
        ref = np.rot90(im.numpy().transpose(1, 2, 0), k).transpose(2, 0, 1)                   # HWC rotated counter-clockwise, like skimage
        np.testing.assert_array_equal(got, ref)
    assert frames.rotate90(im, 1).shape == (3, 7, 5) and frames.rotate90(im, 1).is_contiguous()


@pytest.mark.gpu
def test_gpu_decode_matches_pil_and_prefetcher(tmp_path):
    from PIL import Image
    _write_sequence(str(tmp_path), 3)
    rot = {"A": 0, "B": 1, "C": 3, "Z": 2}
    rot_of = lambda paths: [rot[os.path.basename(p)[0]] for p in paths]
    files_of = lambda t: frames.list_frame_files(str(tmp_path), "seq", t + 1, blacklist=["skip"])
    pf = frames.FramePrefetcher(files_of, rot_of, "cuda:0", num_frames=3)
    for t in range(3):
        ims = pf.get(t)
        paths = files_of(t)
        assert len(ims) == 4
        for im, p in zip(ims, paths):
            ref = np.array(Image.open(p)).astype(np.float64) / 255.0                           # what get_dataset starts from
            ref = np.rot90(ref, rot[os.path.basename(p)[0]]).transpose(2, 0, 1)
            assert im.dtype == torch.float32 and im.is_cuda and tuple(im.shape) == ref.shape
            err = np.abs(im.cpu().numpy().astype(np.float64) - ref)
            if p.endswith(".png"):
                assert err.max() < 1e-7                                                        # lossless: only the float32 rounding
            else:
                assert err.max() <= 6.0 / 255 + 1e-7 and err.mean() < 1.0 / 255, (err.max() * 255, err.mean() * 255)   # measured: 5.0 / 0.74
    assert not pf._pending                                                                      # nothing queued past the last frame
