"""Run bench.py over a grid of (views, blend_px) on one GPU and print one compact line per point."""
import itertools
import json
import subprocess
import sys

views = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "3,6,12").split(",")]
pxs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4").split(",")]
extra = sys.argv[3:]
for v, px in itertools.product(views, pxs):
    out = subprocess.run([sys.executable, "bench.py", "--views", str(v), "--blend-px", str(px), "--steps", "20", "--warmup", "3",
                          "--no-cpu-baseline", "--no-e2e", *extra], capture_output=True, text=True).stdout.strip().splitlines()
    try:
        d = json.loads(out[-1])
        st = d["roofline"]["all_stage_ms_per_launch"]
        print(f"views {v:3d} px {px} step {d['ms_per_step']:.3f} ms  " + " ".join(f"{k}={x:.3f}" for k, x in st.items()), flush=True)
    except Exception as e:
        print("views", v, "px", px, "FAILED", e, out[-3:], flush=True)
