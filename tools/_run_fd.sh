#!/bin/bash
mkdir -p gpurun_out
(
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fir_down or downconv" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_generator.py -x -q -m gpu 2>&1 | tail -3
for m in 0 1 0 1; do
  echo "== N3D_FIR_STREAM=$m (both FIR kernels)"
  N3D_FIR_STREAM=$m timeout 600 python bench.py --steps 30 --warmup 5 --no-other-configs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:^fir_down" -s 21 -c 7 --csv --log-file gpurun_out/firdown.csv python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
grep -c fir_down gpurun_out/firdown.csv
) > gpurun_out/r2_fd.log 2>&1
tail -40 gpurun_out/r2_fd.log
