for cfg in c2 c4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --config $cfg > gpurun_out/r2_bench_n2_$cfg.json 2> gpurun_out/r2_bench_n2_$cfg.err
tail -2 gpurun_out/r2_bench_n2_$cfg.err
done
python bench.py --config c4 --steps 2 --warmup 1 > gpurun_out/r2_bench_n1_c4.json 2>/dev/null
