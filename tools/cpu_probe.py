import sys, time, os, torch
sys.path.insert(0, '.')
from next3d_b200 import config, weights
from oracle import generator as og
cfg = config.full_config(512)
sd = weights.make_state_dict(cfg, seed=0)
z, cc, c, v = weights.demo_inputs(cfg, 1)
uc, uf = weights.sampler_noise(cfg, 1)
for th in (8, 16, 32, 64, 128):
    torch.set_num_threads(th)
    with torch.no_grad():
        ws = og.mapping(sd, cfg, z, cc, 0.7, 14)
        og.synthesis(sd, cfg, ws, c, v, uc, uf)
        t = time.perf_counter(); og.synthesis(sd, cfg, ws, c, v, uc, uf); dt = time.perf_counter() - t
    print(th, 'threads', round(dt, 2), 's/img', flush=True)
