#!/usr/bin/env bash
# Multi-GPU session: bash tools/gpu_session_scale.sh <N> <tag> [what...]   (run under gpurun --gpus N)
set -u
N="$1"; TAG="$2"; shift 2
OUT=gpurun_out
mkdir -p "$OUT"
run() {
    local secs="$1" name="$2"; shift 2
    echo "== $name (limit ${secs}s)" | tee -a "$OUT/${TAG}_session.log"
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=10 "$secs" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
    echo "   exit $? after $((SECONDS - t0))s" | tee -a "$OUT/${TAG}_session.log"
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for what in "$@"; do
    case "$what" in
        check)   run 240 "mg_check" $TR --master-port 29521 tools/multigpu_check.py config2 2 ;;
        config2) run 240 "bench_config2" $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 ;;
        sepwait) MVOC_EXCHANGE_WAIT=separate run 240 "bench_config2_sepwait" $TR --master-port 29528 bench.py --gpus $N --steps 10 --warmup 3 ;;
        nccl)    MVOC_EXCHANGE=nccl run 240 "bench_config2_nccl" $TR --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 ;;
        gather)  MVOC_FP_GATHER_MAX_PIXELS=256 run 240 "bench_config2_gather256" $TR --master-port 29524 bench.py --gpus $N --steps 10 --warmup 3 ;;
        config3) run 240 "bench_config3" $TR --master-port 29525 bench.py --gpus $N --workload config3 --steps 6 --warmup 3 ;;
        config5) run 400 "bench_config5" $TR --master-port 29526 bench.py --gpus $N --workload config5 --steps 3 --warmup 2 --no-cpu-baseline ;;
        check5)  run 400 "mg_check5" $TR --master-port 29527 tools/multigpu_check.py config5 1 ;;
    esac
done
echo "== done" | tee -a "$OUT/${TAG}_session.log"
