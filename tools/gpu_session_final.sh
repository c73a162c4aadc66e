#!/usr/bin/env bash
# Final single-GPU session: what the driver runs at round end (tests, smoke, default bench, reference arm).
set -u
TAG="${1:-r02final}"
OUT=gpurun_out
mkdir -p "$OUT"
run() {
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=10 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
run 900 pytest_gpu python -m pytest tests -m gpu -x -q
run 300 smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run 600 bench_default python bench.py
run 400 bench_reference python bench.py --impl reference
echo "== done" | tee -a "$OUT/${TAG}_session.log"
