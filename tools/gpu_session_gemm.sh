#!/usr/bin/env bash
# GEMM session: gemm_check correctness (all sections, both CTA modes) + timings, isolated kernel timings, kernel tests, a short bench.
set -u
TAG="${1:-r02k}"
OUT=gpurun_out
mkdir -p "$OUT"
run() {
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=10 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
for v in 1 0; do
    for sec in linear conv tconv geglu; do
        run 90 "gemm_${sec}_v${v}" tools/gemm_check "$sec" "$v"
    done
done
run 150 gemm_time_v1 tools/gemm_check time 1
run 120 kernel_times python tools/gpu_diag.py time
run 300 pytest_kernels python -m pytest tests/test_kernels_gpu.py tests/test_properties_gpu.py -m gpu -x -q
run 200 bench_n1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
echo "== done" | tee -a "$OUT/${TAG}_session.log"
