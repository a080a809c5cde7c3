#!/bin/bash
mkdir -p gpurun_out
(
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fir_up or upconv" 2>&1 | tail -5
FIR_ONLY=1 timeout 600 python tools/bench_layers.py fir 2>&1 | grep -v Warn
for m in 0 1 0 1; do
  echo "== N3D_FIR_STREAM=$m"
  N3D_FIR_STREAM=$m timeout 600 python bench.py --steps 30 --warmup 5 --no-other-configs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done
) > gpurun_out/r2_fir.log 2>&1
tail -40 gpurun_out/r2_fir.log
