#!/usr/bin/env bash
# Tensor-parallel measurements for BASELINE config 4 (Qwen3-32B, TP=N) with each all-reduce form, plus a
# per-kernel timeline of rank 0. Costs N x box time: every item has a tight timeout.
#
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/validate_tp8.sh'            # N=8
#   gpurun --gpus 2 --timeout 600 -- 'N=2 bash tools/validate_tp8.sh quick'  # plumbing check on 2 GPUs
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
S=gpurun_out/r02_tp${N}_summary.log
run() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a $S
  timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1
  echo "rc=$? $(grep -E '^\{' "gpurun_out/${name}.log" | tail -n 1 | cut -c1-260)" | tee -a $S
}
B="bench.py --gpus $N --tp --model qwen3-32b --no-cpu-baseline --steps 32 --warmup 4 --mem-util 0.85"
case "${1:-full}" in
quick)   # does the NVLS form survive CUDA-graph decode under torchrun? does the driver's N>1 line work?
  run r02_tp${N}_06b_nvls   200 env B200_ALLREDUCE=nvls B200_BENCH_DEBUG=150 $TR --master-port 29611 bench.py --gpus $N --tp --model qwen3-0.6b --no-cpu-baseline --steps 16 --warmup 3
  run r02_tp${N}_06b_nvlsf  200 env B200_ALLREDUCE=nvls_fused B200_BENCH_DEBUG=150 $TR --master-port 29612 bench.py --gpus $N --tp --model qwen3-0.6b --no-cpu-baseline --steps 16 --warmup 3
  run r02_scale_line_n${N}  500 env B200_BENCH_DEBUG=450 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 3
  ;;
full)
  run r02_tp${N}_32b_forms    560 env B200_BENCH_DEBUG=520 $TR --master-port 29601 tools/tp_forms_bench.py --model ${MODEL:-qwen3-32b} --forms ${FORMS:-auto,nvls_fused,nccl,nvls} --budget-s ${BUDGET:-330}
  grep -E '^\{' gpurun_out/r02_tp${N}_32b_forms.log | tee -a $S | cut -c1-400
  ;;
timeline)
  run r02_tp${N}_32b_timeline 240 $TR --master-port 29605 tools/step_timeline.py --model qwen3-32b --tp --mem-util 0.85 --out gpurun_out/r02_timeline_tp${N}_32b.txt
  ;;
esac
cat $S
