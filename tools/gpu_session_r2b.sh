#!/usr/bin/env bash
# Round-2 session B: the tcgen05 GEMM family on the product path — full GPU suite, A/B against cuDNN / cuBLAS, bench.
set -u
TAG="r02b"
OUT=gpurun_out
mkdir -p "$OUT"
run() {
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=10 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
run 60 gemm_conv_v0 tools/gemm_check conv 0
run 60 gemm_conv_v1 tools/gemm_check conv 1
run 600 pytest_gpu python -m pytest tests -m gpu -x -q -s
run 200 dense_diag python tools/gpu_diag.py dense
run 200 bench_tc_v0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
MVOC_GEMM_VARIANT=1 run 200 bench_tc_v1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
echo "== done" | tee -a "$OUT/${TAG}_session.log"
