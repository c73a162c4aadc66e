#!/usr/bin/env bash
# Round-2 session C: GEMM epilogue rewrite (per-warp pipelines) — correctness, timings, A/B, bench.
set -u
TAG="r02c"
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
run 200 dense_diag python tools/gpu_diag.py dense
run 200 bench_tc python bench.py --steps 10 --warmup 3 --no-cpu-baseline
echo "== done" | tee -a "$OUT/${TAG}_session.log"
