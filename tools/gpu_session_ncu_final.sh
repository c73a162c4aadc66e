set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=r02final
cap() {
    timeout --signal=TERM --kill-after=10 150 ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -f -o "$OUT/${TAG}_ncu_$1" python tools/ncu_kernels.py "$1" > "$OUT/${TAG}_ncu_$1.log" 2>&1
    if [ -f "$OUT/${TAG}_ncu_$1.ncu-rep" ]; then
        python tools/ncu_summary.py "$OUT/${TAG}_ncu_$1.ncu-rep" > "$OUT/${TAG}_ncu_$1_summary.txt" 2>&1
        rm -f "$OUT/${TAG}_ncu_$1.ncu-rep"
    fi
}
cap conv gemm_tc_kernel
cap linear gemm_tc_kernel
cap geglu gemm_tc_kernel
cap tconv gemm_tc_kernel
timeout --signal=TERM --kill-after=20 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "mvoc_timed_region/" --csv --log-file "$OUT/${TAG}_launches_timed_step.csv" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graphs > "$OUT/${TAG}_launches_bench.log" 2>&1
