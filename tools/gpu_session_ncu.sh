#!/usr/bin/env bash
# ncu --set full captures of every kernel family at its benchmark shape + the launch list of one timed step.
set -u
TAG="${1:-r02}"
if [ -x tools/gemm_check ]; then
    echo "== GEMM tile-width experiment" | tee -a "gpurun_out/${TAG}_ncu_session.log"
    mkdir -p gpurun_out
    timeout 120 tools/gemm_check width 1 > "gpurun_out/${TAG}_gemm_width_pairs.txt" 2>&1
    timeout 120 tools/gemm_check width 0 > "gpurun_out/${TAG}_gemm_width_single.txt" 2>&1
fi
OUT=gpurun_out
mkdir -p "$OUT"
cap() {  # cap <name> <kernel regex>
    echo "== ncu $1" | tee -a "$OUT/${TAG}_ncu_session.log"
    timeout --signal=TERM --kill-after=10 150 ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -f \
        -o "$OUT/${TAG}_ncu_$1" python tools/ncu_kernels.py "$1" > "$OUT/${TAG}_ncu_$1.log" 2>&1
    echo "   exit $?" | tee -a "$OUT/${TAG}_ncu_session.log"
    # the reports (tens of MB each with sources) stay on the box: gpurun_out/ carries <= 64 MiB back
    if [ -f "$OUT/${TAG}_ncu_$1.ncu-rep" ]; then
        python tools/ncu_summary.py "$OUT/${TAG}_ncu_$1.ncu-rep" > "$OUT/${TAG}_ncu_$1_summary.txt" 2>&1
        ncu -i "$OUT/${TAG}_ncu_$1.ncu-rep" --page details --csv > "$OUT/${TAG}_ncu_$1_details.csv" 2>/dev/null
        ls -la "$OUT/${TAG}_ncu_$1.ncu-rep" >> "$OUT/${TAG}_ncu_session.log"
        rm -f "$OUT/${TAG}_ncu_$1.ncu-rep"
    fi
}
cap attn attn_fwd_kernel
cap attn_pair attn_fwd_kernel
cap attn_temporal attn_temporal
cap gn gnh_apply
cap gn_t gnh_stats
cap layernorm layernorm_kernel
cap qk_blend qk_blend_kernel
cap feature_blend feature_blend_kernel
cap conv gemm_tc_kernel
cap conv_l2 gemm_tc_kernel
cap linear gemm_tc_kernel
cap geglu gemm_tc_kernel
cap tconv gemm_tc_kernel
echo "== launch list of one timed step (eager)" | tee -a "$OUT/${TAG}_ncu_session.log"
timeout --signal=TERM --kill-after=20 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx \
    --nvtx-include "mvoc_timed_region/" --csv --log-file "$OUT/${TAG}_launches_timed_step.csv" \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graphs > "$OUT/${TAG}_launches_bench.log" 2>&1
echo "   exit $?" | tee -a "$OUT/${TAG}_ncu_session.log"
echo "== N=1 bench, K=5 (for kernels.top)" | tee -a "$OUT/${TAG}_ncu_session.log"
timeout --signal=TERM --kill-after=20 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_bench_n1_k5.log" 2>&1
echo "   exit $?" | tee -a "$OUT/${TAG}_ncu_session.log"
