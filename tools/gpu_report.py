"""Stage-by-stage parity report on the GPU box (does not stop at the first failure).
    python tools/gpu_report.py [--tiny] > gpurun_out/report.txt
"""
import json
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import parity  # noqa: E402
from topo4d_b200 import synth  # noqa: E402


def main():
    tiny = "--tiny" in sys.argv
    print("device:", torch.cuda.get_device_name(0), "| torch", torch.__version__, flush=True)
    cases = [("tiny_precomp", lambda: (synth.random_scene(300, 0), [synth.front_camera(64, 48)], 48, 64, 0, (0.2, 0.5, 0.8)))]
    if not tiny:
        cases += [
            ("cfg1_precomp", lambda: (synth.random_scene(5000, 0), [synth.front_camera(256, 256)], 256, 256, 0, (0, 0, 0))),
            ("cfg1_sh3", lambda: (synth.random_scene(5000, 13, sh_degree=3), [synth.make_camera(synth.look_at((1.5, 0.8, -3.5)), 256, 256, 256.0, 256.0)], 256, 256, 3, (0.1, 0.1, 0.1))),
            ("multi_view3", lambda: (synth.random_scene(3000, 5, sh_degree=2), [synth.make_camera(synth.look_at(e), 160, 128, 200.0, 200.0) for e in ((0, 0, -4.0), (3.0, 1.0, -2.5), (-2.0, -1.5, 3.0))], 128, 160, 2, (0, 0, 0))),
            ("head_1080p_2views", lambda: (synth.head_scene(60000, 0, 3, "topo4d"), synth.ring_cameras(24)[:2], 1080, 1920, 3, (0, 0, 0))),
            ("head_1080p_generic", lambda: (synth.head_scene(60000, 0, 3, "generic"), synth.ring_cameras(24)[5:6], 1080, 1920, 3, (0, 0, 0))),
        ]
    report = {}
    for name, mk in cases:
        t0 = time.time()
        try:
            scene, cams, H, W, deg, bg = mk()
            m = parity.compare(scene, cams, H, W, deg, bg)
            try:
                parity.assert_parity(m, allow_flips=max(2, m["pixels"] // 100000))
                m["verdict"] = "PASS"
            except AssertionError:
                m["verdict"] = "FAIL"
        except Exception:
            m = {"verdict": "ERROR", "trace": traceback.format_exc()}
        m["seconds"] = round(time.time() - t0, 2)
        report[name] = m
        print(name, json.dumps(m, indent=1), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(report, open("gpurun_out/report.json", "w"), indent=1)


if __name__ == "__main__":
    main()
