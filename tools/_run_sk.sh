#!/bin/bash
# split-K A/B: kernel tests, micro-benchmark, whole-step A/B (same box, back to back)
mkdir -p gpurun_out
(
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "splitk or conv" 2>&1 | tail -5
timeout 600 python tools/bench_splitk.py 2>&1 | grep -v Warn
for cfg in "0 16" "1 8" "1 16" "1 32" "0 16" "1 16"; do
  set -- $cfg
  echo "== N3D_SPLITK=$1 MAXRES=$2"
  N3D_SPLITK=$1 N3D_SPLITK_MAXRES=$2 timeout 600 python bench.py --steps 30 --warmup 5 --no-other-configs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['gpu_launches'])"
done
timeout 900 python -m pytest tests/test_gpu_generator.py -x -q -m gpu 2>&1 | tail -5
) > gpurun_out/r2_splitk.log 2>&1
tail -40 gpurun_out/r2_splitk.log
