#!/usr/bin/env bash
# Round-2 session A: bring-up of the tcgen05 GEMM family + config-2 parity tests.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session_r2a.sh'
set -u
TAG="r02a"
OUT=gpurun_out
mkdir -p "$OUT"
run() {  # run <seconds> <name> <command...>
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=10 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > "$OUT/${TAG}_gpu.txt" 2>&1
for v in 0 1; do
    for sec in linear conv tconv geglu; do
        run 90 "gemm_${sec}_v${v}" tools/gemm_check $sec $v
    done
done
run 120 gemm_time_v0 tools/gemm_check time 0
run 120 gemm_time_v1 tools/gemm_check time 1
run 90 split_attn tools/staged_check attn
run 60 split_time tools/staged_check time
ok0=1
for sec in linear conv tconv geglu; do
    grep -q ": 0 failing" "$OUT/${TAG}_gemm_${sec}_v0.log" || ok0=0
done
echo "gemm v0 ok: $ok0" | tee -a "$OUT/${TAG}_session.log"
if [ "$ok0" = "1" ]; then
    run 900 pytest_gpu python -m pytest tests -m gpu -x -q -s
    run 150 dense_diag python tools/gpu_diag.py dense
    run 200 bench_tc python bench.py --steps 10 --warmup 3 --no-cpu-baseline
    MVOC_DENSE=lib run 200 bench_lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline
else
    MVOC_DENSE=lib run 900 pytest_gpu_lib python -m pytest tests -m gpu -x -q -s
    MVOC_DENSE=lib run 200 bench_lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline
fi
run 120 kernel_times python tools/gpu_diag.py time
echo "== done" | tee -a "$OUT/${TAG}_session.log"
