#!/usr/bin/env bash
# ncu --set full captures of every kernel family at its benchmark shape + the launch list of one timed step.
set -u
TAG="${1:-r02}"
OUT=gpurun_out
mkdir -p "$OUT"
cap() {  # cap <name> <kernel regex>
    echo "== ncu $1" | tee -a "$OUT/${TAG}_ncu_session.log"
    timeout --signal=TERM --kill-after=10 150 ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -f \
        -o "$OUT/${TAG}_ncu_$1" python tools/ncu_kernels.py "$1" > "$OUT/${TAG}_ncu_$1.log" 2>&1
    echo "   exit $?" | tee -a "$OUT/${TAG}_ncu_session.log"
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
