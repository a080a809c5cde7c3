#!/bin/bash
# compute-sanitizer over the kernels added late in round 2 (streamed FIR kernels, split-K reduction); appended to profiles/r02_sanitizer.txt
mkdir -p gpurun_out
(
for tool in memcheck synccheck; do
  echo "=================== ${tool}_streamed_fir_and_splitk"
  timeout 200 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "(streamed and (3-128-256 or 4-256-128 or 8-512-64 or 5-512-64)) or splitk" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Error|error" | head -8
done
) > gpurun_out/sanitizer_new.txt 2>&1
cat gpurun_out/sanitizer_new.txt
