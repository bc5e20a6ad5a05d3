#!/usr/bin/env python
"""Benchmark of the DCGAN adversarial training step (G+D iteration of train_gan.py:121-150) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B | --global-batch G] [--dtype bf16|fp32]
                    [--nc 1|3] [--no-check] [--no-rgb] [--no-wgan]

N>1 is launched by the driver as `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`
(one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1]): DCGAN nz=100, ngf=ndf=64, nc=1, batch 512 per GPU, bf16 storage / fp32
accumulate, synthetic uniform[-1,1] images.  NOTE: BASELINE.json says "64x64"; the reference architecture is
hard-wired to 224x224 (dcgan.py:26,84; a 64x64 input raises in Discriminator), so every number is at 224x224.

metric  = training images/s over all ranks.  Default: weak scaling, per-GPU batch fixed at 512 (configs[1] per GPU).
          --global-batch G (BASELINE.json configs[2] uses 4096): strong scaling, per-GPU batch G / N, `"scaling": "strong"`.
--check = (default at N>1; --no-check skips it) before timing, every rank verifies the data-parallel invariants (replicas bit-identical after the steps, gradients
          equal to the sum over ranks) and the line carries `"dp_check": "ok"`; a violation aborts the run.
value   = device-timed (CUDA events, max over ranks), inputs resident in HBM.
e2e     = same metric through DCGANTrainer.step with HOST inputs: every step copies the real batch and the noise
          from pinned host memory and reads the 5 history scalars back.
roofline= the dominant kernel class (conv_gemm_tc_kernel): best case (the D3 conv forward, M=B*196 K=2048 N=256, timed alone with
          CUDA events, L2 flushed) and `class_frac`: the class's 24 launches weighted by their in-step durations (CUPTI pass after
          the timed region; never inside it).
cpu_baseline / --impl reference = the reference's CPU path (oracle/torch_cpu_port.py, stock torch.nn on all host
          threads) on a bounded sample (batch 64 per step) of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'dcgan_train_images_per_sec'
UNIT = 'images/s'
FLOP_PER_IMAGE = {1: 9.169e9, 3: 9.400e9}     # algorithmic conv FLOPs per image per iteration by nc (dead wgrad excluded)


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']), src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(nme)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the step (port), all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import torch
    import torch_cpu_port as port
    # each step = one G+D iteration on a bounded sample of the batch (64 images: BASELINE configs[0]; ~0.7 s on 16 cores), smaller
    # when many steps are asked for, so that the whole run stays within a few minutes
    sample_batch = 64 if args.steps <= 100 else 16
    steps = args.steps
    r = port.time_cpu_steps(batch=sample_batch, steps=steps, warmup=min(args.warmup, 2), nc=args.nc, threads=port.host_threads())
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['images_per_s'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'DCGAN train step nz=100 ngf=ndf=64 nc={args.nc} 224x224 (reference is hard-wired to 224x224, not 64x64)',
                   'per_gpu_batch': args.batch, 'sample': f'each step = one G+D iteration on a {sample_batch}-image sample of the batch'},
        'cpu_baseline': {'value': r['images_per_s'], 'unit': UNIT, 'cores': r['threads'], 'kind': 'port',
                         'sample': f'{steps} iterations x batch {sample_batch} after {min(args.warmup, 2)} warm-up (oracle/torch_cpu_port.py: stock torch.nn CPU '
                                   'path of dcgan.py/train_gan.py; a PORT, /root/reference does not exist on the GPU box)'},
        'e2e': {'value': r['images_per_s'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'host': {'nproc': os.cpu_count(), 'torch_threads': torch.get_num_threads()},
    }
    print(json.dumps(line), flush=True)


def _time_launch(torch, launch, flush, iters):
    """Average CUDA-event duration of one launch on the launching stream, L2 flushed (256 MB write) before every launch."""
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def time_dominant_kernel(pkg, torch, batch, peaks, iters=20):
    """roofline: the kernel class with the largest share of the step is conv_gemm_tc_kernel (fprop / dgrad of D2-D4 and G1-G3,
    25 % of the iteration in profiles/r01_f_step_profile.txt); its representative launch, the D3 forward convolution
    (Conv2d 128->256, k4 s2 p1, 28x28 -> 14x14, BatchNorm statistics fused in the epilogue exactly as the step runs it), is
    timed alone.  `more` adds the other kernel families of the step, each against the roofline that bounds it."""
    import ctypes as C
    L = pkg._lib
    bf = torch.bfloat16
    flush = torch.empty(256 * 1024 * 1024, device='cuda', dtype=torch.uint8)      # > 126 MB L2
    st = L.stream_ptr

    def rnd(shape, dt=bf):
        return torch.randn(shape, device='cuda').to(dt)
    n, ci, h, co, k = batch, 128, 28, 256, 4
    x, w, y = rnd((n, h, h, ci)), torch.randn((co, ci, k, k), device='cuda') * 0.02, torch.empty((n, h // 2, h // 2, co), device='cuda', dtype=bf)
    cv = L.Conv(k, 2, 1, L.ALGO_TCGEN05)                                           # fails loudly if the tensor-core path is absent
    wp = torch.empty(w.numel(), device='cuda', dtype=bf)
    L.call('b200gan_pack_conv_weight', L.ptr(w), co, ci, k, 0, L.ptr(wp), st())
    sums = torch.zeros(2 * co, device='cuda', dtype=torch.float64)
    fz = L.fuse(bn_sums=sums)
    ms = _time_launch(torch, lambda: L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(L.view_nhwc(x)), L.ptr(w), L.ptr(wp),
                                            C.byref(L.view_nhwc(y)), C.byref(fz), st()), flush, iters)
    flops = 2.0 * n * (h // 2) ** 2 * co * (16 * ci)
    ach = flops / (ms * 1e-3) / 1e12
    more = []
    # weight gradient of the same layer (tensor bound)
    dy, dw, ws = rnd((n, h // 2, h // 2, co)), torch.zeros_like(w), torch.zeros_like(w)
    m2 = _time_launch(torch, lambda: L.call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(L.view_nhwc(x)), C.byref(L.view_nhwc(dy)), L.ptr(dw),
                                            L.ptr(ws), None, st()), flush, iters)
    more.append({'kernel': 'conv_wgrad_tc_kernel<128,256,2,6> + wgrad_finalize_kernel (D3 weight gradient)', 'bound': 'tensor',
                 'achieved': flops / (m2 * 1e-3) / 1e12, 'peak': peaks['tf_burst'], 'unit': 'TFLOP/s', 'ms_per_launch': m2})
    # BatchNorm backward apply on D1's tensor (HBM bound: read dz, read y, write dy)
    yb, dz = rnd((n, 56, 56, 64)), rnd((n, 56, 56, 64))
    vec = [torch.rand(64, device='cuda') + 0.5 for _ in range(5)]
    bs = torch.zeros(128, device='cuda', dtype=torch.float64)
    m3 = _time_launch(torch, lambda: L.call('b200gan_bn_act_bwd_apply', C.byref(L.view_nhwc(dz)), C.byref(L.view_nhwc(yb)), None, L.ptr(vec[0]),
                                            L.ptr(vec[1]), L.ptr(vec[2]), L.ptr(vec[3]), L.ptr(vec[4]), L.ptr(bs), n * 56 * 56, L.ACT_NONE, 0.2,
                                            C.byref(L.view_nhwc(dz)), None, None, st()), flush, iters)
    byt = 3.0 * yb.numel() * 2
    more.append({'kernel': 'bn_act_bwd_apply_dense_kernel (D1 tensor, 3 passes)', 'bound': 'hbm', 'achieved': byt / (m3 * 1e-3) / 1e9,
                 'peak': peaks['hbm'], 'unit': 'GB/s', 'ms_per_launch': m3})
    # image-side transposed convolution + Tanh (G5 forward; HBM bound: read the 32-channel tensor, write the image)
    a4, w5, img = rnd((n, 112, 112, 32)), torch.randn((32, 1, 4, 4), device='cuda') * 0.02, torch.empty((n, 224, 224, 1), device='cuda', dtype=bf)
    cva = L.Conv(4, 2, 1, L.ALGO_AUTO)
    ft = L.fuse(out_act=L.ACT_TANH)
    m4 = _time_launch(torch, lambda: L.call('b200gan_convT2d_fprop', C.byref(cva), C.byref(L.view_nhwc(a4)), L.ptr(w5), None,
                                            C.byref(L.view_nhwc(img)), C.byref(ft), st()), flush, iters)
    byt = (a4.numel() + img.numel()) * 2.0
    more.append({'kernel': 'thin_up_tma_kernel<1> (G5 forward + Tanh)', 'bound': 'hbm', 'achieved': byt / (m4 * 1e-3) / 1e9, 'peak': peaks['hbm'],
                 'unit': 'GB/s', 'ms_per_launch': m4})
    # the thin "down" layer D1 (32 -> 64 channels) forward with BatchNorm statistics: halo-tile tcgen05 kernel, HBM bound
    # (read the 32-channel input once, write the 64-channel output once)
    x1, w1 = rnd((n, 112, 112, 32)), torch.randn((64, 32, 4, 4), device='cuda') * 0.02
    y1, wp1, s1 = torch.empty((n, 56, 56, 64), device='cuda', dtype=bf), torch.empty(64 * 32 * 16, device='cuda', dtype=bf), torch.zeros(128, device='cuda', dtype=torch.float64)
    L.call('b200gan_pack_conv_weight', L.ptr(w1), 64, 32, 4, 0, L.ptr(wp1), st())
    f1 = L.fuse(bn_sums=s1)
    m5 = _time_launch(torch, lambda: L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(L.view_nhwc(x1)), L.ptr(w1), L.ptr(wp1),
                                            C.byref(L.view_nhwc(y1)), C.byref(f1), st()), flush, iters)
    byt = (x1.numel() + y1.numel()) * 2.0
    more.append({'kernel': 'conv_down4_tc_kernel<1> (D1 forward + BatchNorm statistics)', 'bound': 'hbm', 'achieved': byt / (m5 * 1e-3) / 1e9,
                 'peak': peaks['hbm'], 'unit': 'GB/s', 'ms_per_launch': m5})
    for r in more:
        r['frac'] = r['achieved'] / r['peak']
    return {'bound': 'tensor', 'kernel': 'conv_gemm_tc_kernel<256,64,4,1>: conv2d_fprop D3 + BatchNorm statistics (M=B*196, K=2048, N=256), bf16 tcgen05, '
                                         'timed alone with CUDA events on the launching stream, L2 flushed',
            'achieved': ach, 'peak': peaks['tf_burst'], 'unit': 'TFLOP/s', 'frac': ach / peaks['tf_burst'],
            'traffic': None, 'traffic_note': 'not measurable inside this run (needs ncu); the ncu --set full capture of this launch is in profiles/ '
                                             '(dram__bytes_read.sum + dram__bytes_write.sum per launch); algorithmic bytes 154.1e6 (input 102.8e6 + output 51.4e6)',
            'ms_per_launch': ms, 'peak_source': peaks['src'] + ' (burst: kernel timed alone)', 'more': more}


def kernel_breakdown(torch, step_fn, steps=2):
    """Per-kernel GPU time of `steps` replayed iterations through CUPTI (torch.profiler), run AFTER the timed region.
    Returns {kernel name: (launches per step, microseconds per step)}."""
    import collections
    import re
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step_fn()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r'\(.*', '', e.name.replace('(anonymous namespace)::', '')).replace('void b200gan::', '').replace('b200gan::', '')
            agg[name[:96]][0] += 1
            agg[name[:96]][1] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
    return {k: (n / steps, t / steps) for k, (n, t) in agg.items()}


# 2*MAC of one k4 s2 p1 layer-operation of D1-D4 / G1-G4 per image (102.76 MMAC each)
LAYER_OP_FLOP = 2 * 102.76e6
# launches per iteration of each tcgen05 kernel family: the generic kernel serves D2-D4 fprop (x3 passes) and
# dgrad (x3), G1-G3 fprop and dgrad -- minus D2 dgrad (x3) and G3 fprop, which run on conv_up4w_tc_kernel; the halo-tile kernels D1 fprop (x3) /
# dgrad (x3), G4 fprop / dgrad; wgrad D1-D4 (x2) + G1-G4
TC_CLASS_LAUNCHES = {'conv_gemm_tc_kernel': 20, 'conv_up4w_tc_kernel': 4, 'conv_up4_tc_kernel': 4, 'conv_down4_tc_kernel': 4, 'conv_wgrad_tc_kernel': 12}


def class_fractions(breakdown, batch, peaks):
    """Tensor-core kernel classes weighted over ALL their launches in the step: FLOPs of the class / its summed in-step duration,
    against the sustained bf16 peak (the kernels run back to back inside a long step)."""
    out = {}
    for cls, launches in TC_CLASS_LAUNCHES.items():
        rows = [(k, v) for k, v in breakdown.items() if k.startswith(cls)]
        n = sum(v[0] for _, v in rows)
        us = sum(v[1] for _, v in rows)
        if not rows or us <= 0:
            continue
        tf = n * LAYER_OP_FLOP * batch / (us * 1e-6) / 1e12      # the launches actually seen (every one is one layer-operation), not the expected count
        out[cls] = {'launches_per_step': n, 'expected_launches': launches, 'us_per_step': us, 'tflops': tf,
                    'frac_of_sustained_peak': tf / peaks['tf_sustained'], 'frac_of_burst_peak': tf / peaks['tf_burst']}
    return out


def dp_check(torch, dist, tr, world, rank, B, nz, nc):
    """Data-parallel invariants, asserted on every rank before the timed region (driver-visible multi-GPU correctness):
      1. after the exchange, every rank holds the SAME gradient arena, and it equals the sum over ranks of the local gradients
         (summed a second time through torch.distributed as an independent path);
      2. after three full steps with the buckets overlapping the backward pass (kernel by kernel, capture, graph replay) the
         weights and Adam moments are bit-identical on all ranks."""
    import gan_enhanced_pneumonia_classifier_b200.trainer as T
    gen = torch.Generator(device='cuda').manual_seed(1234 + rank)
    real = torch.rand((B, nc, 224, 224), device='cuda', generator=gen) * 2 - 1
    noise = torch.randn((B, nz, 1, 1), device='cuda', generator=gen)
    snap = [t.clone() for a in (tr.arenaG, tr.arenaD) for t in (a.param, a.exp_avg, a.exp_avg_sq)]
    bufs = [(b, b.clone()) for net in (tr.netG, tr.netD) for b in net.buffers()]

    def rewind():
        with torch.no_grad():
            for a, k in ((tr.arenaG, 0), (tr.arenaD, 3)):
                a.param.copy_(snap[k]); a.exp_avg.copy_(snap[k + 1]); a.exp_avg_sq.copy_(snap[k + 2])
                a.step_dev.zero_(); a.step = 0
            for b, v in bufs:
                b.copy_(v)
        tr.refresh_packed_weights()

    # (1) the local D gradients of one run (no bucket goes out during the backward pass: overlap=False), then every bucket through
    #     the library's communicator; the same local gradients summed through torch.distributed are the independent answer
    g = tr._segments(real, noise, overlap=False)
    assert next(g) == 'D'
    want = tr.arenaD.grad.clone()
    dist.all_reduce(want)
    tr._exchange('D', overlapped=False)
    torch.cuda.synchronize()
    got = tr.arenaD.grad.clone()
    g.close()
    rewind()
    err = float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))
    assert err < 1e-6, f'rank {rank}: exchanged D gradients differ from the sum over ranks (max rel {err:.3e})'
    ref = got.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, got), f'rank {rank}: gradient arena differs from rank 0 after the exchange'
    # (2) replicas stay bit-identical over steps
    for _ in range(3):
        tr.step(real, noise)
    torch.cuda.synchronize()
    for a in (tr.arenaG, tr.arenaD):
        for t in (a.param, a.exp_avg, a.exp_avg_sq):
            r0 = t.clone()
            dist.broadcast(r0, 0)
            assert torch.equal(r0, t), f'rank {rank}: replica state diverged from rank 0'
    rewind()
    return 'ok'


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gan_enhanced_pneumonia_classifier_b200 as pkg
    from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: there is no CPU fallback for the product path')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    peaks = load_peaks()
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    nz, nc = 100, args.nc
    strong = args.global_batch is not None
    if strong and args.global_batch % world:
        raise SystemExit(f'--global-batch {args.global_batch} is not divisible by {world} ranks')
    B = args.global_batch // world if strong else args.batch

    def make_trainer(nc_):
        torch.manual_seed(0)
        G, D = pkg.Generator(nz, nc_, 64).cuda(), pkg.Discriminator(nc_, 64).cuda()
        G.apply(pkg.weights_init)
        D.apply(pkg.weights_init)
        if world > 1:      # replicas start from rank 0's weights, as DDP would
            for t in list(G.state_dict().values()) + list(D.state_dict().values()):
                dist.broadcast(t, 0)
        return DCGANTrainer(G, D, lr=2e-4, beta1=0.5, dtype=dtype, sync_bn=args.sync_bn)

    tr = make_trainer(nc)
    gen = torch.Generator(device='cuda').manual_seed(1 + rank)
    real = torch.rand((B, nc, 224, 224), device='cuda', generator=gen) * 2 - 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_start = time.perf_counter()

    def log(msg):
        if os.environ.get('B200GAN_BENCH_VERBOSE'):
            print(f'[bench rank {rank} +{time.perf_counter() - t_start:6.1f}s] {msg}', file=sys.stderr, flush=True)

    log('dp check')
    check = dp_check(torch, dist, tr, world, rank, min(B, 64), nz, nc) if (args.check and world > 1) else None
    log('warm-up')

    # ---- device-resident timing: inputs already in HBM (the trainer's static input buffers), fresh on-device noise per step
    #      as in the reference (train_gan.py:132) -------------------------------------------------
    in_real, in_noise = tr.input_buffers(real.shape, torch.float32, (B, nz, 1, 1))
    in_real.copy_(real)
    for _ in range(max(args.warmup, 3)):
        tr.step(in_real, in_noise.normal_(generator=gen))
    barrier()
    log('timed region')
    l0 = tr.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        m = tr.step(in_real, in_noise.normal_(generator=gen))
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = tr.launches - l0
    last = m.cpu().tolist()

    # ---- end to end: host buffers, H2D every step, D2H of the history scalars ----------------------
    # The user-facing call is DCGANTrainer.step(real, noise).  Every step copies THAT step's real batch and noise from pinned
    # host memory (on a copy stream, into one of two staging buffers, so the transfer of step i+1 overlaps the kernels of
    # step i: plain double buffering, what a DataLoader with pin_memory + non_blocking copies gives) and reads the five
    # history scalars back to pinned host memory.  The host batch is bf16: the first convolution rounds the image to bf16
    # for its tensor-core operand anyway (bit-identical results, test_gpu_step.py), so shipping fp32 only doubles the PCIe /
    # host-memory traffic that eight ranks share.
    e2e_dt = torch.bfloat16 if dtype == torch.bfloat16 else torch.float32
    real_h = torch.empty((B, nc, 224, 224), dtype=e2e_dt).pin_memory()
    real_h.copy_(real.cpu())
    noise_h = [torch.empty((B, nz, 1, 1), dtype=torch.float32).pin_memory() for _ in range(2)]
    hist_h = torch.empty(5, dtype=torch.float32).pin_memory()
    static_real, static_noise = tr.input_buffers(real.shape, e2e_dt, (B, nz, 1, 1))
    stage_r = [torch.empty((B, nc, 224, 224), device='cuda', dtype=e2e_dt) for _ in range(2)]
    stage_n = [torch.empty((B, nz, 1, 1), device='cuda') for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    copy_stream = torch.cuda.Stream()

    def enqueue_copy(i):
        k = i % 2
        copied[k].synchronize()                            # host buffer k was last read by the copy issued two iterations ago
        noise_h[k].normal_()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])            # the staging buffer was read by the step two iterations ago
            stage_r[k].copy_(real_h, non_blocking=True)
            stage_n[k].copy_(noise_h[k], non_blocking=True)
            copied[k].record(copy_stream)

    def e2e_loop(n):
        main = torch.cuda.current_stream()
        for k in range(2):
            consumed[k].record(main)
        enqueue_copy(0)
        for i in range(n):
            k = i % 2
            main.wait_event(copied[k])
            static_real.copy_(stage_r[k])
            static_noise.copy_(stage_n[k])
            consumed[k].record(main)
            if i + 1 < n:
                enqueue_copy(i + 1)
            hist_h.copy_(tr.step(static_real, static_noise), non_blocking=True)

    log('end-to-end')
    e2e_loop(3)                     # first call for this input dtype runs kernel by kernel, second captures, third replays
    barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(e2e_steps)
    f1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(f0.elapsed_time(f1), wall_ms)
    t = torch.tensor([ms, ms_e2e], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- after the timed regions (never inside): per-kernel breakdown of replayed steps, the RGB configuration, the CPU baseline
    # (every rank runs the same steps -- they contain collectives --, only rank 0 watches them through CUPTI)
    step_fn = lambda: tr.step(in_real, in_noise.normal_(generator=gen))      # noqa: E731
    log('per-kernel breakdown')
    # One kernel at a time for the per-kernel numbers: inside the timed iteration the side streams (weight gradients, the
    # Generator forward) overlap kernels, which stretches each kernel's own duration without saying anything about the kernel.  The breakdown
    # therefore re-captures the same iteration on ONE stream (all ranks alike: the iteration contains collectives).
    tr.engD.wgrad_stream = tr.engG.wgrad_stream = tr.gfwd_stream = None
    tr._graphs.clear()
    for _ in range(3):
        step_fn()
    if rank == 0:
        breakdown = kernel_breakdown(torch, step_fn)
    else:
        breakdown = None
        for _ in range(2):
            step_fn()
    barrier()
    rgb = None
    log('rgb configuration')
    if nc == 1 and not args.no_rgb and not strong:
        tr.close()
        del tr, stage_r, static_real
        torch.cuda.empty_cache()
        tr3 = make_trainer(3)
        real3 = torch.rand((B, 3, 224, 224), device='cuda', generator=gen) * 2 - 1
        n3 = torch.empty((B, nz, 1, 1), device='cuda')
        for _ in range(3):
            tr3.step(real3, n3.normal_(generator=gen))
        barrier()
        k3 = max(5, min(args.steps, 30))
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(k3):
            tr3.step(real3, n3.normal_(generator=gen))
        g1.record()
        barrier()
        t3 = torch.tensor([g0.elapsed_time(g1)], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        rgb = {'nc': 3, 'note': 'the CLI default (train_gan.py --num-channels 3), the only configuration generate_synthetic.py loads',
               'value': B * world * k3 / (t3.item() * 1e-3), 'unit': UNIT, 'ms_per_step': t3.item() / k3, 'steps': k3}

    # BASELINE.json configs[4]: WGAN-GP (wggan.py / train_wggan.py), critic with gradient-penalty double backward, n_critic = 5, batch 512 per GPU.
    # A parity-test configuration (tests/test_gpu_wgan.py), measured here for the record: one iteration = 5 critic updates + 1 generator update
    # on one batch of real images; images/s counts the real images consumed.  Kernel by kernel (no CUDA graph yet), local BatchNorm statistics.
    wgan = None
    log('wgan-gp configuration')
    extras = world == 1 or args.extras_dp        # the two widened configurations: on one GPU by default, data parallel with --extras-dp

    def timed_ms(fn, k):
        """k calls of fn between two events, barriers on both sides, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            out = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / k], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), out

    if nc == 1 and not args.no_wgan and not strong and extras:
        for name in ('tr3', 'tr'):
            if name in locals() and locals()[name] is not None:
                locals()[name].close()
        tr = tr3 = None
        torch.cuda.empty_cache()
        from gan_enhanced_pneumonia_classifier_b200 import wggan
        from gan_enhanced_pneumonia_classifier_b200.wgan_trainer import WGANGPTrainer
        torch.manual_seed(0)
        wG, wD = wggan.Generator(nz, 1, 64).cuda(), wggan.Discriminator(1, 64).cuda()
        wtr = WGANGPTrainer(wG, wD, critic_iters=5, dtype=dtype)
        wreal = torch.rand((B, 1, 224, 224), device='cuda', generator=gen) * 2 - 1
        wtr.step(wreal)
        torch.cuda.synchronize()
        kw = 3
        wl0 = wtr.launches
        wms, wout = timed_ms(lambda: wtr.step(wreal), kw)
        wgan = {'model': 'WGAN-GP (wggan.py, train_wggan.py:66-93): 5 critic updates with gradient penalty (double backward) + 1 generator update per iteration',
                'nc': 1, 'per_gpu_batch': B, 'n_gpus': world, 'value': B * world / (wms * 1e-3), 'unit': 'real images/s', 'ms_per_iteration': wms, 'steps': kw,
                'gpu_launches_per_iteration': (wtr.launches - wl0) // kw, 'last_losses': [float(v) for v in wout.tolist()]}
        wtr.close()
        del wtr, wG, wD
        torch.cuda.empty_cache()

    # BASELINE.json configs[3]: CGAN (cgan.py / train_cgan.py), CLI-default widths (feature_maps 32, nc=3), batch 1024 per GPU.  A parity-test
    # configuration (tests/test_gpu_cgan.py, tests/test_gpu_perceptual.py), measured for the record: one iteration of train_cgan.py:150-193 without
    # and with the VGG16 perceptual term (random VGG16 weights).  Kernel by kernel, unfused bias / BatchNorm passes, no CUDA graph.
    cg = None
    log('cgan configuration')
    if nc == 1 and not args.no_cgan and not strong and extras:
        from gan_enhanced_pneumonia_classifier_b200 import cgan as cgan_mod
        from gan_enhanced_pneumonia_classifier_b200.cgan_trainer import CGANTrainer
        torch.manual_seed(0)
        cB = 1024
        cG, cD = cgan_mod.Generator(nz, 2, 3, 32).cuda(), cgan_mod.Discriminator(2, 3, 32).cuda()
        ctr = CGANTrainer(cG, cD, perceptual_weight=0.0, dtype=dtype)
        creal = torch.rand((cB, 3, 224, 224), device='cuda', generator=gen) * 2 - 1
        clab = torch.randint(0, 2, (cB,), device='cuda', generator=gen)
        ctr.step(creal, clab)
        torch.cuda.synchronize()
        kc = 5
        cms, cout = timed_ms(lambda: ctr.step(creal, clab), kc)
        cg = {'model': 'CGAN (cgan.py, train_cgan.py:150-193): feature_maps 32, projection discriminator; generator loss = adversarial + '
                       '5 x feature matching (`value`: --no-perceptual) and + 10 x VGG16 perceptual (`with_perceptual`: the full loss of train_cgan.py:191, '
                       'VGG16 on RANDOM weights -- the ImageNet checkpoint cannot be obtained offline; the arithmetic does not depend on the values)',
              'nc': 3, 'per_gpu_batch': cB, 'n_gpus': world, 'value': cB * world / (cms * 1e-3), 'unit': UNIT, 'ms_per_iteration': cms, 'steps': kc,
              'last_history': [float(v) for v in cout.tolist()]}
        ctr.close()
        del ctr
        torch.cuda.empty_cache()
        from gan_enhanced_pneumonia_classifier_b200.perceptual import PerceptualLoss
        vgg = PerceptualLoss('random').cuda()
        ctr = CGANTrainer(cG, cD, perceptual=vgg, perceptual_weight=10.0, dtype=dtype)
        ctr.step(creal, clab)
        torch.cuda.synchronize()
        pms, cout = timed_ms(lambda: ctr.step(creal, clab), kc)
        cg['with_perceptual'] = {'value': cB * world / (pms * 1e-3), 'unit': UNIT, 'ms_per_iteration': pms, 'steps': kc, 'last_history': [float(v) for v in cout.tolist()]}
        ctr.close()
        del ctr, cG, cD, vgg
        torch.cuda.empty_cache()
    log('roofline kernels, cpu baseline')
    if rank == 0:
        value = B * world * args.steps / (ms * 1e-3)
        e2e = B * world * e2e_steps / (ms_e2e * 1e-3)
        roof = time_dominant_kernel(pkg, torch, B, peaks)
        cls = class_fractions(breakdown, B, peaks)
        if 'conv_gemm_tc_kernel' in cls:
            roof['class_frac'] = cls['conv_gemm_tc_kernel']['frac_of_sustained_peak']
            roof['class_frac_of_burst_peak'] = cls['conv_gemm_tc_kernel']['frac_of_burst_peak']
            roof['class_note'] = ('conv_gemm_tc_kernel over all its launches of one iteration (fused epilogues included), FLOPs / summed in-step '
                                  'durations (CUPTI, after the timed region: a short pass, so closer to burst clocks than the timed region) against the SUSTAINED bf16 peak; '
                                  '`class_frac_of_burst_peak` is the same against the burst peak; `frac` above is the best single launch')
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import torch_cpu_port as port
        cpu = None
        if not args.no_cpu_baseline:
            r = port.time_cpu_steps(batch=64, steps=12, warmup=1, nc=nc, threads=port.host_threads())     # ~10 s of CPU work on the box's 16-24 threads
            cpu = {'value': r['images_per_s'], 'unit': UNIT, 'cores': r['threads'], 'kind': 'port',
                   'sample': '12 iterations x batch 64 after 1 warm-up (oracle/torch_cpu_port.py: stock torch.nn CPU path of dcgan.py/train_gan.py; '
                             'a PORT: /root/reference does not exist on the GPU box)'}
        top = sorted(breakdown.items(), key=lambda kv: -kv[1][1])[:12]
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak', 'vs_baseline': None,
            'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': f'DCGAN train step nz=100 ngf=ndf=64 nc={nc} 224x224 (reference is hard-wired to 224x224, not 64x64)',
                       'per_gpu_batch': B, 'global_batch': B * world, 'parallelism': f'dp{world}', 'batchnorm': 'synchronised over ranks (--sync-bn)' if (args.sync_bn and world > 1) else 'local (per-rank) statistics',
                       'gradient_exchange': 'none (1 GPU)' if world == 1 else 'bucketed NCCL all-reduce on the library\'s communicator, captured in '
                                                                              'the iteration graph, overlapped with the backward pass',
                       'l2': 'per-step working set (several GB of activations) far exceeds the 126 MB L2; no flush needed',
                       'algo': os.environ.get('B200GAN_ALGO', 'auto')},
            'model_flops_frac_of_peak': value / world * FLOP_PER_IMAGE[nc] / 1e12 / peaks['tf_sustained'],
            'roofline': roof, 'tensor_core_classes': cls, 'cpu_baseline': cpu,
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': real_h.numel() * real_h.element_size() + noise_h[0].numel() * 4,
                    'd2h_bytes_per_step': 20, 'ms_per_step': ms_e2e / e2e_steps, 'host_dtype': str(e2e_dt).replace('torch.', '')},
            'gpu_launches': launches, 'clocks': clocks, 'last_history': dict(zip(['errD', 'errG', 'D_x', 'D_G_z1', 'D_G_z2'], last)),
            'kernel_time_ms_per_step': sum(v[1] for v in breakdown.values()) / 1e3,
            'kernel_breakdown_schedule': 'single stream (the timed iteration overlaps weight gradients / the Generator forward on side streams)',
            'top_kernels': [{'kernel': k, 'launches_per_step': v[0], 'us_per_step': round(v[1], 1)} for k, v in top],
        }
        more = [c for c in (rgb, wgan, cg) if c is not None]
        if more:
            line['more_configs'] = more
        if check is not None:
            line['dp_check'] = check
        print(json.dumps(line), flush=True)
    if world > 1:
        for t in [x for x in (locals().get('tr'), locals().get('tr3')) if x is not None]:
            t.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200, help='timed iterations (default: ~2 s of work, so that the clocks line shows the power-capped steady state)')
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=512, help='per-GPU batch')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--nc', type=int, default=1, choices=[1, 3], help='image channels: 1 = the benchmark workload (BASELINE.json), 3 = the CLI default')
    ap.add_argument('--global-batch', type=int, default=None, help='strong scaling: total batch over all ranks (BASELINE configs[2]: 4096)')
    ap.add_argument('--check', dest='check', action='store_true', default=True,
                    help='N>1 (default on): assert the data-parallel invariants on every rank before timing; the line carries "dp_check": "ok"')
    ap.add_argument('--no-check', dest='check', action='store_false', help='skip the data-parallel invariants check')
    ap.add_argument('--no-rgb', action='store_true', help='skip the additional nc=3 measurement (more_configs)')
    ap.add_argument('--no-wgan', action='store_true', help='skip the additional WGAN-GP measurement (more_configs, 1 GPU only)')
    ap.add_argument('--sync-bn', action='store_true', help='N>1: synchronised BatchNorm statistics (default: local to each rank)')
    ap.add_argument('--extras-dp', action='store_true', help='N>1: also run the WGAN-GP and CGAN measurements (more_configs) data parallel; default: one GPU only')
    ap.add_argument('--no-cgan', action='store_true', help='skip the additional CGAN measurement (more_configs, 1 GPU only)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
