"""Per-region instruction breakdown of one kernel from an .ncu-rep (source page).
    python tools/ncu_hot.py rep.ncu-rep kernel_regex [min_pct]
Consecutive SASS instructions with (nearly) the same execution count are merged into one region."""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr_i = [i for i, r in enumerate(rows) if "Instructions Executed" in r]
for n, hi in enumerate(hdr_i):
    hdr = rows[hi]
    end = hdr_i[n + 1] - 1 if n + 1 < len(hdr_i) else len(rows)
    iA, iS, iE, iT = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    data = []
    for r in rows[hi + 1:end]:
        try:
            data.append((r[iA], int(r[iS]), int(r[iE]), r[iT]))
        except Exception:
            pass
    tot = sum(d[2] for d in data) or 1
    tots = sum(d[1] for d in data) or 1
    print(rows[hi - 1][:2] if hi > 0 else "", "total warp instructions", tot)
    prev = None; start = 0; acc = 0; accs = 0
    for i, (s, sm, e, t) in enumerate(data + [("", 0, -10**12, "")]):
        if prev is None or abs(e - prev) > 0.03 * max(prev, 1):
            if prev is not None and acc / tot * 100 >= minp:
                print(f"{start:4d}-{i-1:4d} n={i-start:3d} each {prev:10d} total {acc/tot*100:5.1f}% samp {accs/tots*100:5.1f}% thr {data[start][3]:>3s} | {data[start][0].strip()[:56]}")
            start = i; acc = 0; accs = 0; prev = e
        acc += e; accs += sm
