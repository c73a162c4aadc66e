#!/usr/bin/env bash
# Round-2 session D: pair attention + inject entry + fp16, GEMM epilogue v2, ncu captures of the GEMM kernel.
set -u
TAG="r02d"
OUT=gpurun_out
mkdir -p "$OUT"
run() {
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=10 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
for v in 0 1; do
    for sec in linear conv tconv geglu; do
        run 60 "gemm_${sec}_v${v}" tools/gemm_check $sec $v
    done
done
run 600 pytest_gpu python -m pytest tests -m gpu -x -q
run 120 pair_diag python tools/gpu_diag.py pair
run 200 dense_diag python tools/gpu_diag.py dense
run 200 bench_tc python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run 120 ncu_linear ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f \
    -o "$OUT/${TAG}_linear_qkv_l0" tools/gemm_check one linear 1 327680 320 960 0
run 120 ncu_conv ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f \
    -o "$OUT/${TAG}_conv_l0_960" tools/gemm_check one conv 1 80 64 64 960 320
echo "== done" | tee -a "$OUT/${TAG}_session.log"
