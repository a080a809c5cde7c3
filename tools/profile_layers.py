"""Per-launch device times of the tensor-core convolution kernel inside one real generator forward (CUDA events around
every n3d_conv_gemm launch, eager mode).  Usage on the GPU box:  python tools/profile_layers.py [batch] > profiles/....txt"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from next3d_b200 import config, weights
from next3d_b200.triplane_next3d import TriPlaneGenerator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = config.full_config(512)
G = TriPlaneGenerator.from_config(cfg, weights.make_state_dict(cfg, 0), device='cuda')
z, cc, c, v = weights.demo_inputs(cfg, B)
ws = G.mapping(z.cuda(), cc.cuda(), truncation_psi=0.7, truncation_cutoff=14)
eng = G._get_engine()
for _ in range(3):
    G.synthesis(ws, c.cuda(), v.cuda(), noise_mode='const', seed=1)
eng.prof = []
eng.concurrent = False          # per-kernel times: no concurrent branches
G.synthesis(ws, c.cuda(), v.cuda(), noise_mode='const', seed=1)
torch.cuda.synchronize()
rows = [(info, e0.elapsed_time(e1) * 1e3, fl) for kind, e0, e1, fl, info in eng.prof if kind == 'conv_gemm']
tot = sum(r[1] for r in rows)
print(f'batch {B}: {len(rows)} conv_gemm launches (incl. the split-K reduction passes, listed with the layer), {tot/1e3:.2f} ms, {sum(r[2] for r in rows)/tot/1e6:.1f} TFLOP/s algorithmic (x3 executed)')
print(f'{"layer":58s} {"Cin":>5s} {"Cout":>5s} {"M-space":>9s} {"taps":>4s} {"us":>8s} {"TF/s":>7s} {"%":>5s}')
agg = {}
for (name, cin, cout, mh, mw, taps), us, fl in rows:
    k = (name, cin, cout, mh if taps in (9, 1) and True else mh, taps)
    a = agg.setdefault((name, cin, cout), [0, 0.0, 0.0, mh, mw, 0])
    a[0] += 1; a[1] += us; a[2] += fl; a[5] += taps
for (name, cin, cout), (n, us, fl, mh, mw, taps) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{name:58s} {cin:5d} {cout:5d} {mh:4d}x{mw:<4d} {taps:4d} {us:8.1f} {fl/us/1e6:7.1f} {100*us/tot:5.1f}')
