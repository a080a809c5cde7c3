#!/bin/bash
N=${1:-4}
for cfg in c2 c4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --config $cfg > gpurun_out/r2_bench_n${N}_$cfg.json 2> gpurun_out/r2_bench_n${N}_$cfg.err
tail -2 gpurun_out/r2_bench_n${N}_$cfg.err
cat gpurun_out/r2_bench_n${N}_$cfg.json | cut -c1-600
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 --impl reference > gpurun_out/r2_bench_n${N}_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_n${N}_ref.json
