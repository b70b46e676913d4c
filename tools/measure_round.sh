#!/usr/bin/env bash
# One GPU-box call that produces every artefact profiles/ needs for a round tag (default: rXX):
#   gpurun --timeout 900 -- 'bash tools/measure_round.sh r02a'
# then, back in the build container:   python tools/collect_profiles.py r02a
# Writes into gpurun_out/: <tag>_pytest.log, <tag>_bench.json, <tag>_launches.csv (ncu --metrics gpu__time_duration.sum
# --clock-control none, the launch list the bench shares are checked against), <tag>_all.ncu-rep (ncu --set full of every
# rasterizer kernel of one step), <tag>_loss.json (+ --with-loss-ncu: <tag>_loss.ncu-rep), <tag>_dropin.json.
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
(timeout 560 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > "$OUT/${TAG}_pytest.log"
timeout 240 python bench.py --steps 20 --warmup 3 > "$OUT/${TAG}_bench.json" 2> "$OUT/${TAG}_bench.err"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-upstream-style --workloads config2 > "$OUT/${TAG}_l.log" 2>&1
# 4 untimed/warm-up steps x 7 kernels precede the profiled step
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"blend|sort_gather|preprocess|scatter|scan" \
    --launch-skip 28 -c 7 -o "$OUT/${TAG}_all" -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-upstream-style --workloads config2 > "$OUT/${TAG}_n.log" 2>&1
timeout 200 python tools/bench_loss.py --json "$OUT/${TAG}_loss.json" > "$OUT/${TAG}_loss.log" 2>&1
if [ "${2:-}" = "--with-loss-ncu" ]; then
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:"ssim" --launch-skip 4 -c 2 -o "$OUT/${TAG}_loss" -f \
        python tools/profile_loss.py 4 > "$OUT/${TAG}_loss_ncu.log" 2>&1
fi
timeout 200 python tools/train_synthetic.py --frames 3 --iters 200 --graph --json "$OUT/${TAG}_train_graph.json" > "$OUT/${TAG}_train.log" 2>&1
cat "$OUT/${TAG}_pytest.log"
python - "$OUT/${TAG}_bench.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print("bench:", round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "Mpix/s  e2e", round(d["e2e"]["value"]), " frac", round(r["frac"], 3))
    print("stages:", {k: round(v, 4) for k, v in r["all_stage_ms_per_launch"].items()})
except Exception as e:  # noqa: BLE001
    print("bench line unreadable:", e)
PY
