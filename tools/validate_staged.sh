#!/usr/bin/env bash
# One gpurun call that takes every STAGED path over its first run on a B200
# and leaves one log per item under gpurun_out/ . Each item runs under its own timeout so that a
# deadlock in one does not cost the rest (and never leaves the GPU hung for the gpurun limit).
#
#   gpurun --timeout 2400 -- 'bash tools/validate_staged.sh'            # 1 GPU
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/validate_staged.sh tp'  # the 2-GPU items only
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, timeout seconds, command...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/validate_summary.log
  timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1
  echo "rc=$? $(tail -n 1 "gpurun_out/${name}.log" | cut -c1-200)" | tee -a gpurun_out/validate_summary.log
}
if [ "${1:-}" = "tp" ]; then
  run tp_default        600 python -m pytest tests/test_tp_gpu.py -x -q
  run tp_nvls_kernel    300 env B200_NVLS_ALLREDUCE=1 python -m pytest tests/test_tp_gpu.py -q -k nvls
  run tp_nvls_engine    600 env B200_NVLS_ALLREDUCE=all python -m pytest tests/test_tp_gpu.py -q -k tp2_matches_tp1
  run tp_twoshot_engine 600 env B200_TWOSHOT_ALLREDUCE=1 python -m pytest tests/test_tp_gpu.py -q -k tp2_matches_tp1
  exit 0
fi
run moe_emulation       300 python tools/emulate_moe_kernels.py
run moe_gpu             900 env B200_RUN_MOE=1 python -m pytest tests/test_z_moe_gpu.py -q
run moe_30b_big_umma    900 env B200_MOE_BIG_GEMM=umma python bench.py --model qwen3-30b-a3b --no-cpu-baseline --steps 16 --warmup 4 --mem-util 0.85
run umma2_kernels       600 env B200_PREFILL_KERNEL=umma2 python -m pytest tests/test_kernels_gpu.py -q -k prefill
run umma2_engine        900 env B200_PREFILL_KERNEL=umma2 python -m pytest tests/test_engine_gpu.py -q -k "tiny_model or ragged or prefix_cache or llama"
run chunked_prefill     600 env B200_RUN_STAGED=1 python -m pytest tests/test_z_staged_gpu.py -q -k chunked
run kbench_vs_fa2       600 python tools/kbench.py all
run ref_workload_b256    600 python tools/ref_bench_workload.py
run ref_workload_b16     600 python tools/ref_bench_workload.py --block-size 16
run bench_default       600 python bench.py --no-cpu-baseline --steps 64 --warmup 4
run l2prefetch_engine   600 env B200_DECODE_L2_PREFETCH=1 python -m pytest tests/test_engine_gpu.py -q -k "tiny_model or cuda_graph_decode_equals_eager"
run bench_l2prefetch1   600 env B200_DECODE_L2_PREFETCH=1 python bench.py --no-cpu-baseline --steps 64 --warmup 4
run bench_l2prefetch2   600 env B200_DECODE_L2_PREFETCH=2 python bench.py --no-cpu-baseline --steps 64 --warmup 4
run bench_umma2         600 env B200_PREFILL_KERNEL=umma2 python bench.py --no-cpu-baseline --steps 64 --warmup 4
run bench_moe_30b       900 python bench.py --model qwen3-30b-a3b --no-cpu-baseline --steps 32 --warmup 4 --mem-util 0.85
run timeline_moe_30b    900 python tools/step_timeline.py --model qwen3-30b-a3b --mem-util 0.85 --out gpurun_out/timeline_moe_30b.txt
cat gpurun_out/validate_summary.log
