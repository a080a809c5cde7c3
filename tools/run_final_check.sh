#!/bin/bash
R=r02
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests/ -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -2 gpurun_out/bench_$R.err
KR="regex:^(conv_gemm|styles|demod|modulate_split|fir_|splitk_|upsample2d|downsample2d|transform|raster|uv_sample|fill_mouth|mouth_box|resize_aa|blend|render_fused|depth_clamp|mapping)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KR" -s 600 -c 440 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_$R.csv
timeout 300 python tools/profile_layers.py 8 > gpurun_out/layers_b8_$R.txt 2>&1
cut -c1-400 gpurun_out/bench_$R.json
