export N3D_DEBUG_QUANT=0
(timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/debug_render.py 1 16 48 48 1 12 96 96 1 20 48 0 2>&1 | tail -25) > gpurun_out/r2_sanitizer_memcheck_render.txt 2>&1
(timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/debug_render.py 1 16 48 48 1 12 96 96 2>&1 | tail -25) > gpurun_out/r2_sanitizer_racecheck_render.txt 2>&1
(timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke 2>&1 | tail -25) > gpurun_out/r2_sanitizer_memcheck_smoke.txt 2>&1
(timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tools/debug_render.py 1 16 48 48 2>&1 | tail -15) > gpurun_out/r2_sanitizer_synccheck_render.txt 2>&1
tail -6 gpurun_out/r2_sanitizer_*.txt
