#!/usr/bin/env bash
# Round-2 session F (2 GPUs): peer-memory exchange bring-up + multi-rank parity; single-GPU re-checks on GPU 0.
set -u
TAG="r02f"
OUT=gpurun_out
mkdir -p "$OUT"
run() {
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=10 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run 200 mg_reduced2 $TR --master-port 29511 tools/multigpu_check.py reduced2 3
run 300 mg_config2 $TR --master-port 29512 tools/multigpu_check.py config2 2
run 300 bench_n2_peer $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3
MVOC_EXCHANGE=nccl run 300 bench_n2_nccl $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3
run 120 kernel_times python tools/gpu_diag.py time
run 100 geglu_diag env MVOC_DIAG_VARIANTS=1 python tools/gpu_diag.py dense
run 200 bench_n1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
MVOC_GN_FUSED=0 run 200 bench_n1_gn3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
echo "== done" | tee -a "$OUT/${TAG}_session.log"
