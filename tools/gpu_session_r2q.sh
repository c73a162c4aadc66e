#!/usr/bin/env bash
# Round-2 session Q: upsample kernel; full GPU suite + bench.
set -u
TAG="${1:-r02q}"
OUT=gpurun_out
mkdir -p "$OUT"
run() {
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=10 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
run 700 pytest_gpu python -m pytest tests -m gpu -x -q
run 200 bench_n1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
echo "== done" | tee -a "$OUT/${TAG}_session.log"
