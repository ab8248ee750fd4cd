#!/usr/bin/env python
"""Benchmark of the tri-plane render hot path (BASELINE.json metric: ray-samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode fp32|bf16]

A "step" is one ImportanceRenderer.forward over one batch of synthetic input of the BASELINE
config-2 shape: 8 images x 128^2 rays x (48 coarse + 48 importance) samples, 3x32x256^2 planes per
image, random-init OSGDecoder.  One process per GPU (torchrun for N > 1); every rank renders its own
batch of 8 images (weak scaling) and the rendered features / depth / weight sums are all-gathered
over NCCL, as BASELINE.json's north_star describes.  Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_IMG, RES, PLANE_RES, DC, DF = 8, 128, 256, 48, 48
BYTES_PER_SAMPLE = 1536            # 3 planes x 4 taps x 32 channels x 4 B  
METRIC = 'ray-samples/sec'
WORKLOAD = 'config2: batch 8 x 128^2 rays x (48+48) samples, 3x32x256^2 fp32 planes/image'


def ncu_traffic(mode):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` summary of this same command (profiles/); None if there is no capture for this mode."""
    name = {'fp32': 'r01_render_ws_fp32_ncu_full_summary.txt', 'bf16': 'r01_render_ws_bf16_ncu_full_summary.txt'}.get(mode)
    path = os.path.join(ROOT, 'profiles', name) if name else None
    if not path or not os.path.exists(path):
        return None
    tot = 0.0
    for line in open(path):
        for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            if line.startswith(key):
                unit = line[line.index('[') + 1:line.index(']')]
                tot += float(line.split('=')[1]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}[unit]
    return tot or None


def l2_gather_peak():
    """Best random-128-byte-line gather bandwidth measured on this pod's B200 with the working set in L2 (one image's
    planes): tpr_gather_microbench via profiles/gather_roofline.py -> profiles/r01_gather_roofline.json.  None if absent."""
    path = os.path.join(ROOT, 'profiles', 'r01_gather_roofline.json')
    try:
        d = json.load(open(path))
        return max(v for row in d['one_cta_per_sm_25MB_warps_x_lines_in_flight'].values() for v in row.values())
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured'
    return {'hbm_gbs': 6650.0}, 'fallback'


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '20', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        t0, t1 = getattr(self, 't0', 0.0), getattr(self, 't1', float('inf'))
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.03] or [r for _, r in self.rows[-3:]]
        sm = [float(r[0]) for r in rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'reasons': reasons, 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------
# synthetic workload
# ----------------------------------------------------------------------------------------------
def make_inputs(torch, dev, seed, n_img=N_IMG, res=RES):
    cams = importlib.import_module('g-nerf_b200.camera_utils')
    g = torch.Generator(device='cpu').manual_seed(seed)
    planes = torch.randn((n_img, 3, 32, PLANE_RES, PLANE_RES), generator=g, dtype=torch.float32)
    c2w, K = cams.orbit_cameras(n_img)
    return planes, torch.from_numpy(c2w), torch.from_numpy(K)


def make_decoder(torch, pkg, dev, seed):
    torch.manual_seed(seed)
    return pkg.OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).to(dev).requires_grad_(False)


OPTS = {'ray_start': 2.25, 'ray_end': 3.3, 'box_warp': 1, 'depth_resolution': DC, 'depth_resolution_importance': DF,
        'disparity_space_sampling': False, 'clamp_mode': 'softplus'}        # train.py:312-313,328-332


# ----------------------------------------------------------------------------------------------
# CPU baseline (oracle port) -- the only place bench.py executes oracle/
# ----------------------------------------------------------------------------------------------
def cpu_baseline(sample_res, repeats=1):
    """Times the CPU restatement of the reference renderer on a bounded sample of the workload:
    1 image of the same planes/decoder shape, sample_res^2 rays, 48+48 samples."""
    from oracle import triplane_oracle as O
    try:
        from oracle import c_oracle
        have_c = c_oracle.available()
    except Exception:
        have_c = False
    scene = O.synthetic_scene(3, 1, sample_res, PLANE_RES, DC, DF)
    n_samples = sample_res * sample_res * (DC + DF)
    if have_c:
        cores = c_oracle.use_all_cores()
        c_oracle.render(scene, dict(O.FFHQ_OPTIONS))          # warm-up (page-in, thread pool)
        t0 = time.perf_counter()
        for _ in range(repeats):
            c_oracle.render(scene, dict(O.FFHQ_OPTIONS))
        dt = (time.perf_counter() - t0) / repeats
        kind_note = f'C/OpenMP oracle port, {cores} threads'
    else:
        cores = 1
        t0 = time.perf_counter()
        for _ in range(repeats):
            O.render(scene['planes'], scene['dec'], scene['origins'], scene['dirs'], dict(O.FFHQ_OPTIONS),
                     scene['jitter'], scene['u'])
        dt = (time.perf_counter() - t0) / repeats
        kind_note = 'numpy oracle port, 1 thread'
    return {'value': n_samples / dt, 'unit': METRIC, 'cores': cores, 'kind': 'port',
            'sample': f'1 image x {sample_res}^2 rays x (48+48) samples, 3x32x256^2 planes ({kind_note}); {dt:.2f} s/pass'}, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    try:
        from oracle import c_oracle
        fast = c_oracle.available()
    except Exception:
        fast = False
    res = 128 if fast else 32
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(res)
    times = []
    for _ in range(args.steps):
        base, dt = cpu_baseline(res)
        times.append(dt)
    dt = float(np.mean(times))
    value = res * res * (DC + DF) / dt
    base['value'] = value
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': METRIC, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'sample_per_step': base['sample']},
            'cpu_baseline': base,
            'e2e': {'value': value, 'unit': METRIC, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa_node(torch, index):
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off, so that the pinned host buffers of the e2e path
    are allocated next to the GPU's PCIe root (torchrun does not bind ranks).  Returns the node, or None if unknown."""
    try:
        p = torch.cuda.get_device_properties(index)
        dev = f'{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0'
        node = int(open(f'/sys/bus/pci/devices/{dev}/numa_node').read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--mode', default='fp32', choices=['fp32', 'bf16', 'fp32_ffma'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train-step', action='store_true', help='skip the forward+backward timing (N = 1 only)')
    ap.add_argument('--gather', default='peer', choices=['peer', 'nccl'],
                    help='N > 1: peer = the render kernel stores into every GPU\'s gather buffers over NVLink; '
                         'nccl = render, then an in-place all-gather (A/B)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module('g-nerf_b200')
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the renderer has no CPU path')
    # stdout carries the ONE JSON line and nothing else: libraries that print there (NCCL's version banner) go to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, local)       # before any pinned allocation: first touch decides the node
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    warmup = max(args.warmup, 3)

    planes_h, c2w_h, K_h = make_inputs(torch, dev, seed=100 + rank)
    decoder = make_decoder(torch, pkg, dev, seed=0)
    renderer, sampler = pkg.ImportanceRenderer(), pkg.RaySampler()
    opts = dict(OPTS, decoder_precision=args.mode)
    planes = planes_h.to(dev)
    origins, dirs = sampler(c2w_h.to(dev), K_h.to(dev), RES)
    m = RES * RES
    samples_per_step = N_IMG * m * (DC + DF)

    kernel_events = []
    peer = pkg.parallel.PeerGather(N_IMG, m) if world > 1 and args.gather == 'peer' else None

    def step(record=False):
        """The hot path as a user calls it, inputs resident in HBM.  N > 1: every rank renders its batch into
        its slice of EVERY rank's gather buffers (peer stores in the render kernel's epilogue), the depth range is
        all-reduced (which also completes the exchange) and the gathered depths are clamped."""
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            renderer._timing_events = (e0, e1)
            kernel_events.append((e0, e1))
        if world > 1:
            out = pkg.parallel.render_sharded(renderer, planes, decoder, origins, dirs, opts, peer=peer)
        else:
            out = renderer(planes, decoder, origins, dirs, opts)
        renderer._timing_events = None
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler_clk = ClockSampler(local)
    if rank == 0:
        sampler_clk.start()              # nvidia-smi needs ~100 ms to start: begin before the warm-up
    for _ in range(warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step(record=True)
    ev1.record()
    barrier()
    sampler_clk.window(t_wall0, time.time())
    ms = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler_clk.stop() if rank == 0 else None
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))

    # ---- end to end: the forward with HOST buffers (ImportanceRenderer.forward_host -> tpr_render_host): every step
    # copies that step's planes and rays from pinned host memory and reads rgb / depth / weight sums back to the host
    planes_pin, o_pin, d_pin = planes_h.pin_memory(), origins.cpu().pin_memory(), dirs.cpu().pin_memory()
    out_pin = tuple(torch.empty((N_IMG, m, c), dtype=torch.float32).pin_memory() for c in (32, 1, 1))
    h2d = planes_pin.numel() * 4 + o_pin.numel() * 4 + d_pin.numel() * 4
    d2h = sum(t.numel() * 4 for t in out_pin)

    def e2e_step():
        if world == 1:
            renderer.forward_host(planes_pin, decoder, o_pin, d_pin, opts, out=out_pin)
            return
        # N > 1: the same per-image pipeline; the depth clamp needs the all-reduced range, so depth leaves the device last
        renderer.forward_host(planes_pin, decoder, o_pin, d_pin, opts, out=out_pin, defer_depth=True)
        rng = renderer.last_depth_range
        lo, hi = rng[0:1].clone(), rng[1:2].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        renderer.finish_host_depth(torch.cat([lo, hi]))

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(e2e_steps):
        e2e_step()
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1) / e2e_steps

    # ---- training step: forward + backward w.r.t. planes and decoder, same workload
    train = None
    if world == 1 and not args.no_train_step and args.mode != 'fp32_ffma':
        planes_g = planes.detach().clone().requires_grad_(True)
        dec_g = make_decoder(torch, pkg, dev, seed=0).requires_grad_(True)
        ups = (torch.randn(N_IMG, m, 32, device=dev), torch.randn(N_IMG, m, 1, device=dev), torch.randn(N_IMG, m, 1, device=dev))

        def train_step():
            planes_g.grad = None
            for prm in dec_g.parameters():
                prm.grad = None
            torch.autograd.backward(renderer(planes_g, dec_g, origins, dirs, opts), ups)
        for _ in range(3):
            train_step()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_train = max(3, min(args.steps, 10))
        g0.record()
        for _ in range(n_train):
            train_step()
        g1.record()
        barrier()
        tr_ms = g0.elapsed_time(g1) / n_train
        train = {'ms_per_step': tr_ms, 'value': samples_per_step / (tr_ms * 1e-3), 'unit': METRIC, 'steps': n_train,
                 'what': 'ImportanceRenderer.forward + backward (gradients of planes and the four decoder tensors), '
                         'inputs resident in HBM; forward as above, backward = tpr_render_backward'}
        del planes_g, dec_g, ups

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_ms, kern_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, kern_ms = t.tolist()

    if rank == 0:
        pk, pk_kind = peaks()
        value = world * samples_per_step / (ms * 1e-3)
        achieved = samples_per_step * BYTES_PER_SAMPLE / (kern_ms * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': value, 'unit': METRIC, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16-mlp' if args.mode == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'per_gpu_batch': N_IMG, 'rays_per_image': m, 'samples_per_ray': DC + DF,
                       'decoder_precision': args.mode,
                       'parallelism': (f'image-batch sharding x{world}, ' +
                                       ('outputs gathered by the render kernel (NVLink peer stores) + 2-float all-reduce'
                                        if peer is not None else 'NCCL all-gather of outputs')) if world > 1 else 'single GPU',
                       'l2': 'inputs larger than L2: 201 MB planes + 201 MB repack + 50 MB noise per step (126 MB L2)',
                       'step': 'ImportanceRenderer.forward incl. plane repack, decoder pack, both torch.rand draws'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                         'frac': achieved / pk['hbm_gbs'], 'traffic': ncu_traffic(args.mode), 'peak_kind': pk_kind,
                         'traffic_note': 'DRAM bytes per launch (ncu --set full, profiles/): one image\'s 25 MB of planes stays '
                                         'L2-resident, so the 1536 B/sample are L2/L1 traffic and frac can exceed 1',
                         'kernel': ('render_kernel' if args.mode == 'fp32_ffma' else 'render_ws_kernel') + ' (tpr_render: +2 helper launches of ~2 us)', 'kernel_ms': kern_ms,
                         'algorithmic_bytes_per_launch': samples_per_step * BYTES_PER_SAMPLE,
                         'l2_gather_peak': l2_gather_peak(),
                         'frac_of_l2_gather': (achieved / l2_gather_peak()) if l2_gather_peak() else None,
                         'l2_gather_note': 'second ceiling (SURVEY.md section 8(d)): random 128-B line gather out of L2, '
                                           'tpr_gather_microbench on this pod (profiles/r01_gather_roofline.json), GB/s'},
            'e2e': {'value': world * samples_per_step / (e2e_ms * 1e-3), 'unit': METRIC, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': 5 * args.steps,                    # pack_planes, pack_decoder, range_init, render_ws, finish
            'e2e_api': 'ImportanceRenderer.forward_host (tpr_render_host: per-image H2D / repack+render / D2H pipeline)'
            if world == 1 else 'ImportanceRenderer.forward_host(defer_depth) + 2-float all-reduce + finish_host_depth',
            'clocks': clocks,
        }
        if train is not None:
            line['train_step'] = train
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)            # the CPU baseline uses every host core
            line['cpu_baseline'], _ = cpu_baseline(48)
        line['host_numa_node'] = numa
        os.write(real_stdout, (json.dumps(line) + '\n').encode())
    if peer is not None:
        peer.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
