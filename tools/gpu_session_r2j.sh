#!/usr/bin/env bash
# Round-2 session J: attention with per-warp barrier arrivals, config5 on one GPU, full suite, bench.
set -u
TAG="r02j"
OUT=gpurun_out
mkdir -p "$OUT"
run() {
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=10 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
run 100 trace_plain python tools/attn_trace.py 80 5 4096 0
run 120 kernel_times python tools/gpu_diag.py time
run 100 pair_diag python tools/gpu_diag.py pair
run 600 pytest_gpu python -m pytest tests -m gpu -x -q
run 200 bench_n1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run 300 bench_config5_n1 python bench.py --workload config5 --steps 2 --warmup 1 --no-cpu-baseline
echo "== done" | tee -a "$OUT/${TAG}_session.log"
