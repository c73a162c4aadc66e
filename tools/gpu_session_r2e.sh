#!/usr/bin/env bash
# Round-2 session E: fused GroupNorm, share_p on the product path, shared-memory-port experiment for the GEMM kernel.
set -u
TAG="r02e"
OUT=gpurun_out
mkdir -p "$OUT"
run() {
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=10 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
run 90 port_v0 tools/gemm_check port 0
run 90 port_v1 tools/gemm_check port 1
run 600 pytest_gpu python -m pytest tests -m gpu -x -q
run 120 kernel_times python tools/gpu_diag.py time
run 200 bench_tc python bench.py --steps 10 --warmup 3 --no-cpu-baseline
MVOC_GN_FUSED=0 run 200 bench_gn3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
echo "== done" | tee -a "$OUT/${TAG}_session.log"
