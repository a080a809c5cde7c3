#!/bin/bash
mkdir -p gpurun_out
(
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fir_up or upconv" 2>&1 | tail -3
FIR_ONLY=1 timeout 600 python tools/bench_layers.py fir 2>&1 | grep -v "Warn\|convT"
for m in 0 1 1; do
  echo "== N3D_FIR_STREAM=$m"
  N3D_FIR_STREAM=$m timeout 600 python bench.py --steps 30 --warmup 5 --no-other-configs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done
FIR_ONLY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_up_stream -s 1 -c 1 -f -o gpurun_out/prof_firstream python tools/bench_layers.py fir > /dev/null 2>&1
python tools/summarize_ncu.py gpurun_out/prof_firstream.ncu-rep > gpurun_out/ncu_firstream.txt 2>&1
) > gpurun_out/r2_fir.log 2>&1
tail -40 gpurun_out/r2_fir.log
