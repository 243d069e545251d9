#!/usr/bin/env bash
# Round-2 evidence run on ONE GPU: the full -m gpu suite (with the parity figures printed), the
# kernel table against FA2, ncu captures of the prefill kernels, compute-sanitizer logs.
#   gpurun --timeout 2400 -- 'bash tools/profile_r02.sh [tests|kbench|ncu|sanitize ...]'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what=${*:-tests kbench ncu sanitize}
for w in $what; do
case $w in
tests)
  timeout 1500 python -m pytest tests -m gpu -q -s -rs > gpurun_out/r02_gpu_tests.log 2>&1
  echo "tests rc=$? $(tail -n 1 gpurun_out/r02_gpu_tests.log)"
  grep -E "vs oracle|vs reference" gpurun_out/r02_gpu_tests.log > gpurun_out/r02_parity_measured.log
  cat gpurun_out/r02_parity_measured.log ;;
kbench)
  timeout 600 python tools/kbench.py all > gpurun_out/r02_kbench.jsonl 2> gpurun_out/r02_kbench.err
  echo "kbench rc=$?"
  grep -E "prefill|fa2_varlen" gpurun_out/r02_kbench.jsonl ;;
ncu)
  timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:prefill_umma_kernel -s 4 -c 1 -f -o gpurun_out/r02_prefill_umma python tools/kbench.py prefill16 \
    > gpurun_out/r02_ncu_prefill.log 2>&1
  echo "ncu prefill rc=$?" ;;
sanitize)
  for tool in memcheck racecheck; do
    K="decode_ragged or decode_golden or prefill_golden or prefill_ragged or kv_scatter_golden or prefill_paged or prefill_split or qknorm_rope_cache or layers_golden or fp8_gather or fp8_scatter"
    echo "# compute-sanitizer --tool $tool python -m pytest tests/test_kernels_gpu.py tests/test_z_fp8_kv_gpu.py -k \"$K\"" > gpurun_out/r02_sanitizer_$tool.log
    timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 \
      python -m pytest tests/test_kernels_gpu.py tests/test_z_fp8_kv_gpu.py -q -x -k "$K" \
      >> gpurun_out/r02_sanitizer_$tool.log 2>&1
    echo "sanitizer $tool rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/r02_sanitizer_$tool.log | tail -n 2 | tr '\n' ' ')"
  done ;;
esac
done
