#!/usr/bin/env python
"""Benchmark of the hot path: Next3D generator forward (TriPlaneGenerator.synthesis) images/sec on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Workload (BASELINE.json configs[1]): FFHQ-512 generator forward, 64^2 neural render (48 coarse + 48 importance samples
per ray) -> 512^2 SR, batch 8 per GPU, seeded random-init weights (172.8 M parameters), synthetic latents, the demo FLAME
mesh, synthetic all-ones eye mask, noise_mode='const', sampler uniforms from the in-kernel RNG.
A "step" = one synthesis() call on one batch of 8 samples per GPU (weak scaling: batch per GPU fixed).
`value` = whole-job images/s with inputs resident in HBM; `e2e` = the same through the public API with pinned HOST
inputs (H2D inside the timed region) and the final images read back to the host (D2H inside the timed region).
`--impl reference` times the CPU oracle port (the reference itself is Python that cannot travel to the GPU box; the oracle
is pinned bit-exactly against it, tests/test_oracle_vs_reference.py) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = 'FFHQ-512 generator forward (TriPlaneGenerator.synthesis): 64^2 neural render -> 512^2 SR, 48+48 depth samples, '
METRIC = 'generator images/sec at 512^2 (64^2 neural render, 48+48 samples/ray)'
UNIT = 'img/s'


_JSON_OUT = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print to file descriptor 1 behind Python's back (NCCL's version banner,
    for one): keep a private duplicate of the real stdout for the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def _traffic(kernel):
    """Measured DRAM bytes per launch of `kernel` (dram__bytes_read + dram__bytes_write, ncu capture of one forward at the bench
    batch size, summarised by tools/dram_table.py into profiles/): (bytes, source) or (None, None)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r01_dram_traffic.json')
    try:
        with open(path) as f:
            k = json.load(f)['kernels'][kernel]
        return k['dram_bytes_per_launch'], 'profiles/r01_dram_traffic.json (ncu, %d launches of one batch-8 forward)' % k['launches']
    except (OSError, KeyError, ValueError):
        return None, None


def _peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(hbm=float(p['hbm_gbs']), tf=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])), src='measured (MEASURED_PEAKS.json, sustained)')
    except Exception:
        return dict(hbm=6650.0, tf=1400.0, src='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20', '-i', str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [x.strip() for x in line.split(',')])

    def stop(self, window=None):
        """Median SM clock / throttle reasons of the samples that arrived inside `window` = (t0, t1) perf_counter times of the
        timed region; if nvidia-smi delivered none in there (the region is a few hundred ms), of all samples taken under the same
        load (timed region + end-to-end loop), and says which."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows, scope = self.rows, 'timed region + e2e loop (same load)'
        if window is not None:
            inside = [r for r in self.rows if window[0] <= r[0] <= window[1]]
            if inside:
                rows, scope = inside, 'timed region'
        sm, mx, reasons = [], [], set()
        for r in rows:
            r = r[1:]
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm), 'window': scope}


def _cpu_threads():
    """Threads for the CPU arm.  Measured on the GPU box's 128-core host (tools/cpu_probe.py, batch 1): 8 threads 2.30 s/img,
    16 -> 1.98, 32 -> 2.41, 64 -> 3.75, 128 -> 101.9 (oversubscription of many small convolutions): use the fastest setting."""
    return int(os.environ.get('N3D_CPU_THREADS', min(os.cpu_count() or 1, 16)))


def run_reference(args):
    """CPU arm: the oracle port of the reference's CPU path on the host cores, one image per step."""
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from next3d_b200 import config, weights
    from oracle import generator as og
    cores = _cpu_threads()
    torch.set_num_threads(cores)
    cfg = config.full_config(512)
    sd = weights.make_state_dict(cfg, seed=0)
    z, c_cond, c_cam, v = weights.demo_inputs(cfg, 1, seed=0)
    u_c, u_f = weights.sampler_noise(cfg, 1, seed=0)
    with torch.no_grad():
        ws = og.mapping(sd, cfg, z, c_cond, 0.7, 14)
        for _ in range(args.warmup):
            og.synthesis(sd, cfg, ws, c_cam, v, u_c, u_f)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            og.synthesis(sd, cfg, ws, c_cam, v, u_c, u_f)
        dt = time.perf_counter() - t0
    val = args.steps / dt
    sample = f'{args.steps} steps x 1 image (batch 1) of the same generator/config, fp32 torch CPU ops, {cores} threads (of {os.cpu_count()} host cores; more threads are slower, see bench.py::_cpu_threads)'
    _emit({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD + f'batch {args.batch} per GPU', 'global_batch': args.batch,
                   'sample': 'each step = 1 image of that workload (bounded sample: the CPU path takes ~2 s per image), host cores only'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})


def cpu_baseline_sample():
    """Bounded CPU sample for the N=1 line: 1 warm-up + 3 images, batch 1, all host threads."""
    import torch
    from next3d_b200 import config, weights
    from oracle import generator as og
    cores = _cpu_threads()
    torch.set_num_threads(cores)
    cfg = config.full_config(512)
    sd = weights.make_state_dict(cfg, seed=0)
    z, c_cond, c_cam, v = weights.demo_inputs(cfg, 1, seed=0)
    u_c, u_f = weights.sampler_noise(cfg, 1, seed=0)
    times = []
    with torch.no_grad():
        ws = og.mapping(sd, cfg, z, c_cond, 0.7, 14)
        og.synthesis(sd, cfg, ws, c_cam, v, u_c, u_f)
        for _ in range(3):
            t0 = time.perf_counter()
            og.synthesis(sd, cfg, ws, c_cam, v, u_c, u_f)
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {'value': 1.0 / med, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'1 warm-up + median of 3 synthesis() calls, batch 1, same weights/config, CPU oracle port (fp32 torch ops), {cores} of {os.cpu_count()} host threads (fastest setting)'}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from next3d_b200 import config, weights, distributed as D
    from next3d_b200.triplane_next3d import TriPlaneGenerator

    rank, world, local = D.init_from_env()
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    B = args.batch
    cfg = config.full_config(512)
    sd = weights.make_state_dict(cfg, seed=0)
    G = TriPlaneGenerator.from_config(cfg, sd, device=dev)
    G.use_cuda_graph = not args.no_graph
    del sd
    # global sample ids: rank r owns [r*B, (r+1)*B)
    z, c_cond, c_cam, v = weights.demo_inputs(cfg, B * world, seed=0)
    a, b = D.shard_range(B * world, rank, world)
    z, c_cond, c_cam, v = z[a:b], c_cond[a:b], c_cam[a:b], v[a:b]
    with torch.no_grad():
        ws_host = G.mapping(z.to(dev), c_cond.to(dev), truncation_psi=0.7, truncation_cutoff=14).cpu().pin_memory()
    c_host, v_host = c_cam.pin_memory(), v.pin_memory()
    ws_d, c_d, v_d = ws_host.to(dev), c_host.to(dev), v_host.to(dev)
    eng = G._get_engine()
    side = torch.cuda.Stream(dev)

    def step_device(i):
        if world > 1:                                   # the static graph outputs must not be overwritten while the gather reads them
            torch.cuda.current_stream(dev).wait_stream(side)
        out = G.synthesis(ws_d, c_d, v_d, noise_mode='const', seed=1000 + i)
        if world > 1:                                   # the one collective of the path: gather the images on rank 0
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                D.gather_images(out['image'], dst=0)
        return out

    copy_s = torch.cuda.Stream(dev)
    stage, drained = [None, None], [None, None]

    def step_e2e(i, host_out):
        """One end-to-end step the way a serving loop would run it: H2D of the inputs, synthesis, and the D2H of the images on a
        copy stream from a ping-pong staging buffer, so that the 25 MB read-back of step i overlaps the compute of step i+1
        (the graph's static output buffer is free again after an 8 us device copy)."""
        main = torch.cuda.current_stream(dev)
        w = ws_host.to(dev, non_blocking=True)
        c = c_host.to(dev, non_blocking=True)
        vv = v_host.to(dev, non_blocking=True)
        out = G.synthesis(w, c, vv, noise_mode='const', seed=2000 + i)
        img = out['image']
        if world > 1:
            img = D.gather_images(img, dst=0)
        if img is not None:
            k = i & 1
            if drained[k] is not None:
                main.wait_event(drained[k])                 # the D2H of step i-2 has left this staging buffer
            if stage[k] is None:
                stage[k] = torch.empty_like(img)
            stage[k].copy_(img)
            copy_s.wait_stream(main)
            with torch.cuda.stream(copy_s):
                host_out[: img.shape[0]].copy_(stage[k], non_blocking=True)
                drained[k] = torch.cuda.Event()
                drained[k].record(copy_s)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        for i in range(args.warmup):
            step_device(i)
        barrier()
        launches_per_step = eng.launches
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_begin = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            step_device(i)
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(side)
        e1.record()
        barrier()
        t_end = time.perf_counter()
        dt = D.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)

        # ---- end to end through the public API: pinned host inputs -> synthesis -> images on the host
        host_out = torch.empty(B * world if rank == 0 else B, 3, 512, 512).pin_memory()
        for i in range(min(args.warmup, 3)):
            step_e2e(i, host_out)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for i in range(args.steps):
            step_e2e(i, host_out)
        torch.cuda.current_stream(dev).wait_stream(copy_s)   # the last read-back is inside the timed region
        e3.record()
        barrier()
        dt_e2e = D.max_over_ranks(e2.elapsed_time(e3) / 1e3, dev)
        clocks = sampler.stop((t_begin, t_end)) if sampler else None

        # ---- per-kernel device times (CUDA events around every launch of the two graded kernels), one extra step
        roof = roof_r = None
        if rank == 0:
            eng.prof = []
            G.use_cuda_graph = False                         # events around every launch need the eager path ...
            eng.concurrent = False                           # ... and one stream: a kernel sharing the SMs with another branch is not its own time
            G.synthesis(ws_d, c_d, v_d, noise_mode='const', seed=5)
            summ = eng.profile_summary()
            eng.prof = None
            pk = _peaks()
            n_g, ms_g, fl_g = summ['conv_gemm']
            ach = fl_g / (ms_g * 1e-3) / 1e12
            tr_g, tr_src = _traffic('conv_gemm_kernel') if B == 8 else (None, None)
            tr_r, _ = _traffic('render_kernel') if B == 8 else (None, None)
            roof = {'kernel': 'conv_gemm_kernel (tcgen05 implicit GEMM, all conv layers)', 'bound': 'tensor', 'achieved': ach, 'peak': pk['tf'],
                    'unit': 'TFLOP/s', 'frac': ach / pk['tf'], 'traffic': tr_g, 'traffic_unit': 'DRAM bytes per launch (mean)', 'traffic_source': tr_src,
                    'peak_source': pk['src'], 'launches_per_step': n_g,
                    'ms_per_step': ms_g, 'share_of_step': ms_g / (1e3 * dt / args.steps),
                    'note': 'algorithmic FLOPs (one product per MAC: %.1f GFLOP/img); the bf16x3 scheme executes 3x that on the tensor pipe' % (fl_g / B / 1e9)}
            n_r, ms_r, _ = summ['render_rays']
            M, D_ = cfg.neural_rendering_resolution ** 2, 96
            bytes_img = 3 * 32 * 256 * 256 * 4 + M * 6 * 4 + M * 34 * 4            # planes once + rays + rgb32/depth/wsum out (RNG in kernel)
            ach_r = bytes_img * B / (ms_r * 1e-3) / 1e9
            roof_r = {'kernel': 'render_kernel (fused ray sampler + tri-plane fetch + MLP + compositing)', 'bound': 'hbm', 'achieved': ach_r,
                      'peak': pk['hbm'], 'unit': 'GB/s', 'frac': ach_r / pk['hbm'], 'traffic': tr_r, 'algorithmic_bytes_per_launch': bytes_img * B, 'ms_per_step': ms_r,
                      'share_of_step': ms_r / (1e3 * dt / args.steps), 'flops_per_image': M * D_ * 8320}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    imgs = B * world * args.steps
    line = {
        'metric': METRIC, 'value': imgs / dt, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16x3 (hi/lo split operands, fp32 accumulate) for convs; f32 elsewhere', 'data': 'synthetic',
        'config': {'workload': WORKLOAD + f'batch {B} per GPU', 'global_batch': B * world, 'parallelism': f'batch-sharded x{world}, one gather of images',
                   'launch': 'eager' if args.no_graph else 'one CUDA graph replay per step',
                   'l2': 'no explicit flush: per-step working set (0.7 GB packed weights + >4 GB activations) >> 126 MB L2'},
        'clocks': clocks,
        'e2e': {'value': imgs / dt_e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(ws_host.numel() * 4 + c_host.numel() * 4 + v_host.numel() * 4),
                'd2h_bytes_per_step': int(B * 3 * 512 * 512 * 4), 'ms_per_step': 1e3 * dt_e2e / args.steps},
        'gpu_launches': launches_per_step * args.steps,
        'roofline': roof, 'roofline_renderer': roof_r,
    }
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline_sample()
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=8, help='samples per GPU per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch kernels eagerly instead of replaying the captured CUDA graph')
    args = ap.parse_args()
    _guard_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
