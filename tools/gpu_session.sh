#!/usr/bin/env bash
# One gpurun call that collects everything a round needs from a single-GPU box, each stage under its own
# timeout so that a hang costs minutes, not the budget (round 1 lost ~98 GPU-minutes to two multi-rank runs
# that hung at exit under a 600 s limit).  Usage, from the build container:
#
#   gpurun --timeout 3000 -- 'bash tools/gpu_session.sh r02'   # ~45 GPU-minutes worst case
#
# Output: gpurun_out/<tag>_*  (copy what should be judged into profiles/).
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
run() {  # run <seconds> <name> <command...>
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=20 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > "$OUT/${TAG}_gpu.txt" 2>&1
run 420 pytest_gpu   python -m pytest tests -m gpu -x -q
run 60  smoke        python -c "import __graft_entry__ as g; g.smoke()"
run 240 bench_k10    python bench.py --steps 10 --warmup 3
grep -h '^{' "$OUT/${TAG}_bench_k10.log" | tail -1 > "$OUT/${TAG}_bench_k10.json"
run 120 kernel_times python tools/gpu_diag.py time
# ncu, dominant kernel: one full-set capture of the l0 self-attention launch
run 240 ncu_attn     ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 2 -c 1 \
                         -f -o "$OUT/${TAG}_attn_l0" python tools/attn_profile.py
# launch list of ONE timed step (NVTX range pushed by bench.py), eager launches so every kernel is listed
# (round 1: 150 s covered 45 % of one step under ncu; a full step needs ~6 minutes)
run 780 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none --nvtx \
                         --nvtx-include "mvoc_timed_region/" --csv --log-file "$OUT/${TAG}_launches_timed_step.csv" \
                         python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graphs
# ---- options that were staged without hardware: numerics first, then A/B bench lines (K=10 each)
for sec in conv geglu attn time; do   # standalone binary, one section per process (a trap only loses its own section)
    run 150 staged_check_$sec tools/staged_check $sec
done
MVOC_STAGED=1 run 300 staged_tests  python -m pytest tests/test_staged.py -x -q
MVOC_GN_SLAB_MB=24 run 150 bench_gnslab24 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
MVOC_GN_SLAB_MB=48 run 150 bench_gnslab48 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
MVOC_STAGED=1 run 240 staged_times  python tools/gpu_diag.py staged
if grep -q "passed" "$OUT/${TAG}_staged_tests.log" && ! grep -q "failed" "$OUT/${TAG}_staged_tests.log"; then
    MVOC_STAGED=1 run 420 staged_pytest_gpu python -m pytest tests -m gpu -x -q
    MVOC_STAGED=1 run 150 bench_staged python bench.py --steps 10 --warmup 3 --no-cpu-baseline
fi
echo "== done" | tee -a "$OUT/${TAG}_session.log"
