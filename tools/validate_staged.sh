#!/usr/bin/env bash
# First hardware runs of every STAGED path, one log per item under
# gpurun_out/ . Each item runs under its own timeout so that a deadlock in one does not cost the
# rest (and never leaves the GPU hung until gpurun's own limit).
#
#   gpurun --timeout 900  -- 'bash tools/validate_staged.sh moe'       # ~6 min   sparse-MoE kernels + tiny engine
#   gpurun --timeout 900  -- 'bash tools/validate_staged.sh prefill'   # ~8 min   umma2 prefill, chunked prefill
#   gpurun --timeout 1200 -- 'bash tools/validate_staged.sh bench'     # ~12 min  bench lines: default / umma2 / L2 prefetch, FA2 comparators
#   gpurun --timeout 1500 -- 'bash tools/validate_staged.sh moe30b'    # ~12 min  Qwen3-30B-A3B on one GPU (decode + big-GEMM prefill)
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/validate_staged.sh tp'   # 2-GPU items (NVLS / two-shot all-reduce)
#   (no argument = moe + prefill + bench)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SUMMARY=gpurun_out/validate_summary.log
run() {  # name, timeout seconds, command...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$SUMMARY"
  timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1
  echo "rc=$? $(tail -n 1 "gpurun_out/${name}.log" | cut -c1-300)" | tee -a "$SUMMARY"
}
group_moe() {
  run moe_emulation       300 python tools/emulate_moe_kernels.py
  run moe_diag_check      600 python tools/moe_diag.py check
  run moe_gpu             900 env B200_RUN_MOE=1 python -m pytest tests/test_z_moe_gpu.py -q
  run moe_diag_bench      600 python tools/moe_diag.py bench
}
group_prefill() {
  run umma2_kernels       600 env B200_PREFILL_KERNEL=umma2 python -m pytest tests/test_kernels_gpu.py -q -k prefill
  run umma2_engine        900 env B200_PREFILL_KERNEL=umma2 python -m pytest tests/test_engine_gpu.py -q -k "tiny_model or ragged or prefix_cache or llama"
  run sharded_sampling    300 env B200_RUN_STAGED=1 python -m pytest tests/test_z_staged_gpu.py -q -k sharded_sampling_equals
  run chunked_prefill     600 env B200_RUN_STAGED=1 python -m pytest tests/test_z_staged_gpu.py -q -k chunked
}
group_bench() {
  run kbench_vs_fa2       600 python tools/kbench.py all
  run bench_default       600 python bench.py --no-cpu-baseline --steps 64 --warmup 4
  run l2prefetch_engine   600 env B200_DECODE_L2_PREFETCH=1 python -m pytest tests/test_engine_gpu.py -q -k "tiny_model or cuda_graph_decode_equals_eager"
  run bench_l2prefetch1   600 env B200_DECODE_L2_PREFETCH=1 python bench.py --no-cpu-baseline --steps 64 --warmup 4
  run bench_l2prefetch2   600 env B200_DECODE_L2_PREFETCH=2 python bench.py --no-cpu-baseline --steps 64 --warmup 4
  run fused_gate_up_eng   600 env B200_FUSED_GATE_UP=1 python -m pytest tests/test_engine_gpu.py -q -k "tiny_model or cuda_graph_decode_equals_eager"
  run bench_fused_gate_up 600 env B200_FUSED_GATE_UP=1 python bench.py --no-cpu-baseline --steps 64 --warmup 4
  run bench_umma2         600 env B200_PREFILL_KERNEL=umma2 python bench.py --no-cpu-baseline --steps 64 --warmup 4
  run ref_workload_b256   600 python tools/ref_bench_workload.py
  run ref_workload_b16    600 python tools/ref_bench_workload.py --block-size 16
}
group_moe30b() {
  run bench_moe_30b       900 python bench.py --model qwen3-30b-a3b --no-cpu-baseline --steps 32 --warmup 4 --mem-util 0.85
  run moe_30b_big_umma    900 env B200_MOE_BIG_GEMM=umma python bench.py --model qwen3-30b-a3b --no-cpu-baseline --steps 16 --warmup 4 --mem-util 0.85
  run timeline_moe_30b    900 python tools/step_timeline.py --model qwen3-30b-a3b --mem-util 0.85 --out gpurun_out/timeline_moe_30b.txt
}
group_tp() {
  run tp_default        600 python -m pytest tests/test_tp_gpu.py -x -q
  run tp_nvls_kernel    300 env B200_NVLS_ALLREDUCE=1 python -m pytest tests/test_tp_gpu.py -q -k nvls
  run tp_nvls_engine    600 env B200_NVLS_ALLREDUCE=all python -m pytest tests/test_tp_gpu.py -q -k tp2_matches_tp1
  run tp_nvls_fused_eng 600 env B200_NVLS_ALLREDUCE=fused-all python -m pytest tests/test_tp_gpu.py -q -k tp2_matches_tp1
  run tp_sharded_sample 600 env B200_SHARDED_SAMPLING=1 python -m pytest tests/test_tp_gpu.py -q -k tp2_matches_tp1
  run tp_twoshot_engine 600 env B200_TWOSHOT_ALLREDUCE=1 python -m pytest tests/test_tp_gpu.py -q -k tp2_matches_tp1
}
case "${1:-default}" in
  moe) group_moe ;;
  prefill) group_prefill ;;
  bench) group_bench ;;
  moe30b) group_moe30b ;;
  tp) group_tp ;;
  default) group_moe; group_prefill; group_bench ;;
  *) echo "unknown group $1"; exit 2 ;;
esac
cat "$SUMMARY"
