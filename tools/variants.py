"""Bench every experiment build topo4d_b200/_build/libtopo4d_b200_<tag>.so (made by `python -m topo4d_b200.build --tag
<tag> -D...`) plus the default library on one GPU; one compact line per variant.
    python tools/variants.py [views] [extra bench.py args...]"""
import glob
import json
import os
import subprocess
import sys

views = sys.argv[1] if len(sys.argv) > 1 else "24"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = [""] + sorted(glob.glob(os.path.join(root, "topo4d_b200", "_build", "libtopo4d_b200_*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["TOPO4D_B200_LIB"] = lib
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--views", views, "--steps", "20", "--warmup", "3",
                          "--no-cpu-baseline", "--no-e2e", "--no-upstream-style", "--workloads", "config2", *sys.argv[2:]], capture_output=True, text=True, env=env, cwd=root)
    name = os.path.basename(lib)[len("libtopo4d_b200_"):-3] if lib else "default"
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        st = d["roofline"]["all_stage_ms_per_launch"]
        print(f"{name:12s} views {views} step {d['ms_per_step']:.3f} ms  " + " ".join(f"{k}={x:.3f}" for k, x in st.items()), flush=True)
    except Exception as e:
        print(name, "FAILED", e, out.stderr[-300:], flush=True)
