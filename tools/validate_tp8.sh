#!/usr/bin/env bash
# Tensor-parallel measurements for BASELINE config 4 (Qwen3-32B, TP=8) with every all-reduce form,
# plus per-kernel timelines of rank 0. Costs 8x box time: keep it to one call.
#
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/validate_tp8.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/validate_tp_summary.log
  timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1
  echo "rc=$? $(tail -n 1 "gpurun_out/${name}.log" | cut -c1-300)" | tee -a gpurun_out/validate_tp_summary.log
}
B="bench.py --gpus $N --tp --model qwen3-32b --no-cpu-baseline --steps 32 --warmup 4 --mem-util 0.85"
run tp${N}_32b_nccl      420 env B200_PEER_ALLREDUCE=0 $TR --master-port 29601 $B
run tp${N}_32b_default   420 $TR --master-port 29602 $B
run tp${N}_32b_twoshot   420 env B200_TWOSHOT_ALLREDUCE=1 $TR --master-port 29603 $B
run tp${N}_32b_nvls      420 env B200_NVLS_ALLREDUCE=1 $TR --master-port 29604 $B
run tp${N}_32b_nvlsfused 420 env B200_NVLS_ALLREDUCE=fused $TR --master-port 29607 $B
run tp${N}_32b_nvls_shs  420 env B200_NVLS_ALLREDUCE=1 B200_SHARDED_SAMPLING=1 $TR --master-port 29608 $B
run tp${N}_32b_nvls_tl   420 env B200_NVLS_ALLREDUCE=1 $TR --master-port 29605 tools/step_timeline.py --model qwen3-32b --tp --mem-util 0.85 --out gpurun_out/timeline_tp${N}_32b_nvls.txt
run tp${N}_moe_30b       420 $TR --master-port 29606 bench.py --gpus $N --tp --model qwen3-30b-a3b --no-cpu-baseline --steps 32 --warmup 4 --mem-util 0.85
cat gpurun_out/validate_tp_summary.log
