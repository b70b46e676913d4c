"""Turn the scratch artefacts of `tools/measure_round.sh <tag>` (gpurun_out/) into the tracked summaries under profiles/:
    python tools/collect_profiles.py r02a
  <tag>_bench_1gpu.json, <tag>_launches.csv, <tag>_all_kernels_ncu.txt (tools/ncu_summary.py), <tag>_traffic.json (per-launch
  dram bytes; point bench.py's traffic file at it), <tag>_blend_{fwd,bwd}_hot_regions.txt (tools/ncu_hot.py),
  <tag>_loss_bench.json, <tag>_train_graph_1gpu.json, and a per-step kernel-share table printed from the launch list."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def _src_hash():
    """Content hash of the kernel sources the capture was taken with: bench.py only quotes a traffic file whose hash equals
    the current tree's, so the figure cannot silently go stale when a kernel changes."""
    sys.path.insert(0, ROOT)
    from topo4d_b200 import build
    return build._source_hash()


def main():
    tag = sys.argv[1]
    src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    for a, b in ((f"{tag}_bench.json", f"{tag}_bench_1gpu.json"), (f"{tag}_launches.csv", f"{tag}_launches.csv"),
                 (f"{tag}_loss.json", f"{tag}_loss_bench.json"), (f"{tag}_train_graph.json", f"{tag}_train_graph_1gpu.json")):
        if os.path.exists(os.path.join(src, a)):
            shutil.copy(os.path.join(src, a), os.path.join(dst, b))
    rep = os.path.join(src, f"{tag}_all.ncu-rep")
    if os.path.exists(rep):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, os.path.join(dst, f"{tag}_all_kernels_ncu.txt")],
                       stdout=subprocess.DEVNULL)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        h, u = rows[0], rows[1]
        ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        traffic = {}
        for r in rows[2:]:
            name = r[h.index("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0].replace("_mma_kernel", "_kernel")
            traffic.setdefault(name, int(float(r[ir]) * UNIT[u[ir]] + float(r[iw]) * UNIT[u[iw]]))
        json.dump({"source": f"ncu --set full --clock-control none capture summarised in profiles/{tag}_all_kernels_ncu.txt "
                             f"(gpurun_out/{tag}_all.ncu-rep); command: python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline",
                   "workload": {"views_per_launch": 24, "width": 1920, "height": 1080, "gaussians": 60000, "sh_degree": 3, "opacity": "topo4d"},
                   "kernel_src_hash": _src_hash(), "dram_bytes_per_launch": traffic}, open(os.path.join(dst, f"{tag}_traffic.json"), "w"), indent=1)
        for k in ("fwd", "bwd"):
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "blend_" + k, "1.0"],
                                 capture_output=True, text=True).stdout.splitlines()
            half = out[:max(1, len(out) // 2)] if len(out) > 2 and out[0] == out[len(out) // 2] else out
            open(os.path.join(dst, f"{tag}_blend_{k}_hot_regions.txt"), "w").write("\n".join(half) + "\n")
    ll = os.path.join(src, f"{tag}_launches.csv")
    if os.path.exists(ll):
        rows = list(csv.reader(open(ll)))
        hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
        h = rows[hi]
        ik, iv = h.index("Kernel Name"), h.index("Metric Value")
        data = [(r[ik].split("(")[0].split("::")[-1].split("<")[0], float(r[iv].replace(",", "")) / 1e3) for r in rows[hi + 1:] if len(r) > iv]
        last = max(i for i, (k, _) in enumerate(data) if k == "preprocess_kernel")
        prev = max(i for i, (k, _) in enumerate(data[:last]) if k == "preprocess_kernel")
        step = data[prev:last]
        tot = sum(v for _, v in step)
        print("launch list, one step (us):", {k: round(v, 1) for k, v in step}, "sum", round(tot, 1))
        print("shares (%):", {k: round(100 * v / tot, 1) for k, v in step})
    print("bench.py picks the newest profiles/*_traffic.json whose kernel_src_hash matches the tree")


if __name__ == "__main__":
    main()
