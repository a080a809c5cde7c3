#!/bin/bash
mkdir -p gpurun_out
(
FIR_ONLY=1 FIR_SWEEP=1 timeout 600 python tools/bench_layers.py fir 2>&1 | grep -v Warn
FIR_ONLY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_up_stream -s 1 -c 1 -f -o gpurun_out/prof_firstream python tools/bench_layers.py fir > /dev/null 2>&1
python tools/summarize_ncu.py gpurun_out/prof_firstream.ncu-rep > gpurun_out/ncu_firstream.txt 2>&1
) > gpurun_out/r2_fir2.log 2>&1
tail -120 gpurun_out/r2_fir2.log
