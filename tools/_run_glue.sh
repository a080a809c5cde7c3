#!/bin/bash
R=r02
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k "regex:^(styles|demod|modulate_split|fir_|splitk_|upsample2d|downsample2d|transform|raster|uv_sample|fill_mouth|mouth_box|resize_aa|blend|depth_clamp)" -s 294 -c 98 -f -o gpurun_out/prof_glue_$R python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_glue.log 2>&1
python tools/summarize_ncu.py gpurun_out/prof_glue_$R.ncu-rep longest > gpurun_out/ncu_glue_$R.txt 2>&1; ls -la gpurun_out/prof_glue_$R.ncu-rep; rm -f gpurun_out/prof_glue_$R.ncu-rep
timeout 300 ncu --set full --clock-control none -k "regex:^(mapping|interp_rows)" -c 4 -f -o gpurun_out/prof_map_$R python -m pytest tests/test_gpu_generator.py tests/test_gpu_drivers.py -q -m gpu -k "mapping or interpolate" > /dev/null 2>&1
python tools/summarize_ncu.py gpurun_out/prof_map_$R.ncu-rep longest >> gpurun_out/ncu_glue_$R.txt 2>&1; rm -f gpurun_out/prof_map_$R.ncu-rep
head -3 gpurun_out/ncu_glue_$R.txt | cut -c1-600; grep -c "=== launch" gpurun_out/ncu_glue_$R.txt
