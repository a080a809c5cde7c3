#!/bin/bash
# Round-end measurement pass on the GPU box (one gpurun call): tests, bench lines, ncu launch list + DRAM bytes of one forward,
# ncu --set full captures of every kernel family.  Outputs land in gpurun_out/; tools/launch_table.py, tools/dram_table.py and
# tools/summarize_ncu.py turn them into the files committed under profiles/.
R=${1:-r02}
if [ "$2" != "notests" ]; then rm -f gpurun_out/parity_report.jsonl; timeout 1500 python -m pytest tests/ -m gpu -q 2>&1 | tail -3; fi
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -2 gpurun_out/bench_$R.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$R.json 2>/dev/null
python tools/profile_layers.py 8 > gpurun_out/layers_b8_$R.txt 2>&1
KR="regex:^(conv_gemm|styles|demod|modulate_split|fir_|splitk_|upsample2d|downsample2d|transform|raster|uv_sample|fill_mouth|mouth_box|resize_aa|blend|render_fused|depth_clamp|mapping)"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KR" -s 600 -c 440 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_$R.csv
if [ "$3" != "noreps" ]; then   # kernels unchanged since the last capture: skip with "noreps"
N3D_BENCH_GRID=0 ncu --set full --clock-control none --import-source on -k regex:render_fused -s 3 -c 1 -f -o gpurun_out/prof_render_$R python tools/bench_render.py c2 > /dev/null 2>&1
N3D_BENCH_GRID=0 ncu --set full --clock-control none --import-source on -k regex:render_fused -s 3 -c 1 -f -o gpurun_out/prof_render_c3_$R python tools/bench_render.py c3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:points_fused -s 2 -c 1 -f -o gpurun_out/prof_points_$R python tools/bench_render.py c2 > /dev/null 2>&1
fi
# every non-GEMM, non-renderer kernel of ONE forward (98 launches after 3 warm-up forwards), summarised as the longest launch per kernel
ncu --set full --clock-control none -k "regex:^(styles|demod|modulate_split|fir_|splitk_|upsample2d|downsample2d|transform|raster|uv_sample|fill_mouth|mouth_box|resize_aa|blend|depth_clamp)" -s 294 -c 98 -f -o gpurun_out/prof_glue_$R python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
python tools/summarize_ncu.py gpurun_out/prof_glue_$R.ncu-rep longest > gpurun_out/ncu_glue_$R.txt 2>&1; rm -f gpurun_out/prof_glue_$R.ncu-rep      # gpurun_out/ must stay below 64 MiB
ncu --set full --clock-control none -k "regex:(upfirdn2d_kernel|bias_act_kernel|flrelu)" -c 8 -f -o gpurun_out/prof_ops_$R python -m pytest tests/test_gpu_ops_api.py -q -m gpu > /dev/null 2>&1
python tools/summarize_ncu.py gpurun_out/prof_ops_$R.ncu-rep > gpurun_out/ncu_ops_$R.txt 2>&1; rm -f gpurun_out/prof_ops_$R.ncu-rep
if [ "$3" != "noreps" ]; then
cap() { ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c 1 -f -o gpurun_out/prof_$3_$R python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1; }
cap conv_gemm 395 convsr
cap conv_gemm 314 convT
for k in render render_c3 points convsr convT; do python tools/summarize_ncu.py gpurun_out/prof_${k}_$R.ncu-rep > gpurun_out/ncu_${k}_$R.txt 2>&1; done
fi
ls -la gpurun_out/*_$R.ncu-rep; du -sh gpurun_out
