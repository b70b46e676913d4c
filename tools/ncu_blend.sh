#!/usr/bin/env bash
# Counter snapshot of the blend kernels of one 24-view step (a handful of metrics, not --set full: ~20 s):
#   gpurun --timeout 600 -- 'bash tools/ncu_blend.sh TAG'
TAG="${1:-q}"
OUT=gpurun_out
mkdir -p "$OUT"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_lsu.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_mio_throttle.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_not_selected.ratio,smsp__average_warp_latency_issue_stalled_branch_resolving.ratio,smsp__average_warp_latency_issue_stalled_lg_throttle.ratio,smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio
timeout 300 ncu --metrics $M --clock-control none -k regex:"blend" --launch-skip 8 -c 2 --csv --log-file "$OUT/${TAG}_blend_counters.csv" \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-upstream-style --workloads config2 ${2:-} > "$OUT/${TAG}_blend_counters.log" 2>&1
python - "$OUT/${TAG}_blend_counters.csv" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
if rows:
    H = rows[0]
    ki, mi, vi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
    cur = None
    for r in rows[1:]:
        if r[ki] != cur:
            cur = r[ki]; print(cur[:60])
        print("   %-90s %s" % (r[mi], r[vi]))
PY
