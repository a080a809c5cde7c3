#!/bin/bash
# Round-end measurement pass on the GPU box (one gpurun call): tests, bench line, ncu launch list + DRAM bytes of one forward,
# ncu --set full captures of the graded kernels, per-layer conv table.  Outputs land in gpurun_out/; tools/launch_table.py,
# tools/dram_table.py and tools/summarize_ncu.py turn them into the files committed under profiles/.
R=${1:-r01}
timeout 900 python -m pytest tests/ -m gpu -q --timeout 600 2>&1 | tail -3
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -2 gpurun_out/bench_$R.err
python tools/profile_layers.py 8 > gpurun_out/layers_b8_$R.txt 2>&1
KR="regex:^(conv_gemm|styles|demod|modulate_split|fir_|upsample2d|downsample2d|transform|raster|uv_sample|fill_mouth|mouth_box|resize_aa|blend|render|depth_clamp)"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KR" -s 555 -c 185 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_$R.csv
cap() { ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c 1 -f -o gpurun_out/prof_$3_$R python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; }
cap conv_gemm 395 convsr
cap conv_gemm 314 convT
cap fir_up_epilogue 91 firup
cap render_kernel 3 render
ls -la gpurun_out/*_$R.ncu-rep
