#!/usr/bin/env python
"""Benchmark of the CGG decoder-head hot path (BASELINE.json metric: decoder-head images/s at
1024x1024).  One "step" = one pass of the path (Mask2FormerHeadOpen.forward after the pixel
decoder, open_set/models/mask2former_head.py:787-849) over one batch of synthetic pixel-decoder
outputs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--batch B_per_gpu] [--precision bf16|fp32] [--queries Q]

N>1 is launched by torch.distributed.run, one rank per GPU; images shard by batch, no data-path
collective, so "scaling" is weak.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'decoder_head_images_per_sec_1024'
UNIT = 'images/s'
H = W = 1024
C = 256
FFN = 2048
D_L = 768
NCLS1 = 49
LAYERS = 9


def flops_per_image(Q, height=H, width=W):
    """Dense algorithmic FLOPs of the path per image."""
    HW4 = (height // 4) * (width // 4)
    Ks = [(height // s) * (width // s) for s in (32, 16, 8)]
    f = (LAYERS + 1) * (2 * Q * C * HW4 + 2 * Q * C * (3 * C + D_L + NCLS1))
    for i in range(LAYERS):
        K = Ks[i % 3]
        f += 4 * Q * C * C + 4 * K * C * C + 4 * Q * K * C
    f += LAYERS * (8 * Q * C * C + 4 * Q * Q * C)
    f += LAYERS * (4 * Q * C * FFN)
    return float(f)


def einsum_flops_per_launch(Q, batch):
    return 2.0 * Q * C * (H // 4) * (W // 4) * batch


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], tflops=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    tflops_burst=d['bf16_tflops'], source='MEASURED_PEAKS.json')
    return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1590.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None),
                    samples=len(sm), reasons=sorted(reasons))


def reduce_max(value, world, device):
    """MAX over ranks of a per-rank scalar (the timing rule: slowest rank defines the step)."""
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_value(world, batch_per_gpu, steps, ms):
    """images/s of the whole job: every rank processed batch_per_gpu images per step (weak scaling,
    batch-sharded, no data-path collective -- SURVEY.md section 8e)."""
    return world * batch_per_gpu * steps / (ms * 1e-3)


def dist_env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


# ------------------------------------------------------------------------------ CPU legs
def cpu_reference_step(sd, mf, mems):
    from oracle import cgg_oracle as O
    with torch.no_grad():
        return O.decoder_forward(sd, mf, mems)


def cpu_baseline(Q, threads, repeats=3):
    """The oracle (a port of the reference's fp32 PyTorch path) timed on the host cores on a
    bounded sample: one 1024x1024 image per pass."""
    from cgg_b200 import synth
    torch.set_num_threads(threads)
    sd = synth.make_params(seed=0, num_queries=Q)
    mf, mems = synth.make_inputs(0, 1, H, W)
    cpu_reference_step(sd, mf, mems)
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu_reference_step(sd, mf, mems)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return dict(value=1.0 / ts[len(ts) // 2], unit=UNIT, cores=threads, kind='port',
                sample='1 image 1024x1024 per pass, Q=%d, fp32, median of %d passes after 1 warm-up' % (Q, repeats))


def run_reference_arm(args):
    """--impl reference: the reference's CPU path (oracle port; the Python reference itself cannot
    travel to the GPU box) with all host threads, each step a bounded sample of the same workload."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from cgg_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    Q = args.queries
    sample_b = 1
    sd = synth.make_params(seed=0, num_queries=Q)
    mf, mems = synth.make_inputs(0, sample_b, H, W)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(sd, mf, mems)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sd, mf, mems)
    dt = time.perf_counter() - t0
    val = sample_b * args.steps / dt
    line = dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=min(args.warmup, 1),
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', impl='reference',
                config=dict(workload='configs[1]: COCO-OVIS instance decoder head, Q=%d, 9 layers, 256-d, 1024x1024' % Q,
                            batch_per_step=sample_b, note='CPU path, bounded sample of 1 image per step'),
                cpu_baseline=dict(value=val, unit=UNIT, cores=threads, kind='port',
                                  sample='%d image(s) 1024x1024 per step, fp32 oracle port of the reference path' % sample_b),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------- our arm
def run_b200_arm(args):
    import torch.distributed as dist
    from cgg_b200 import synth, lib as clib
    from cgg_b200.head import build_head_from_state_dict

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the B200 path has no CPU fallback')
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    Q, B = args.queries, args.batch
    dt_in = torch.bfloat16 if args.precision == 'bf16' else torch.float32
    sd = synth.make_params(seed=0, num_queries=Q)
    head = build_head_from_state_dict(sd, Q, NCLS1, args.precision, dev, cuda_graph=not args.no_graph,
                                      final_mask_only=args.final_mask_only)
    # synthetic pixel-decoder outputs, per-rank seed; kept in PINNED host memory for the e2e leg
    mf_h, mems_h = synth.make_inputs(rank, B, H, W, dtype=dt_in)
    mf_h = mf_h.pin_memory()
    mems_h = [m.pin_memory() for m in mems_h]
    mf_d = mf_h.to(dev, non_blocking=True)
    mems_d = [m.to(dev, non_blocking=True) for m in mems_h]
    lib = clib.load()

    # `value` keeps args.in_flight batches in flight: one head (own workspace, own CUDA graph) and one stream per
    # slot, steps issued round-robin.  The layer chain of a step is latency-bound, so the
    # chain of one batch fills the SMs the other leaves idle; every step is still a full forward of B images.
    n_fly = max(1, args.in_flight)
    heads = [head] + [build_head_from_state_dict(sd, Q, NCLS1, args.precision, dev, cuda_graph=not args.no_graph,
                                                 final_mask_only=args.final_mask_only)
                      for _ in range(n_fly - 1)]
    fly_in = [(mf_d, mems_d)] + [(mf_d.clone(), [m.clone() for m in mems_d]) for _ in range(n_fly - 1)]
    fly_streams = [torch.cuda.Stream() for _ in range(n_fly)]
    step_no = [0]

    def step_resident():
        i = step_no[0] % n_fly
        step_no[0] += 1
        with torch.cuda.stream(fly_streams[i]):
            return heads[i].decoder_forward(*fly_in[i])

    def fork():
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        for st in fly_streams:
            st.wait_event(ev)

    def join():
        for st in fly_streams:
            torch.cuda.current_stream().wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fork()
    for _ in range(max(args.warmup, 3) * n_fly):
        out = step_resident()
    join()
    barrier()
    # single-batch latency of one step (one stream, nothing else in flight), reported next to the throughput
    l0_, l1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0_.record()
    for _ in range(5):
        head.decoder_forward(mf_d, mems_d)
    l1_.record()
    torch.cuda.synchronize()
    latency_ms = l0_.elapsed_time(l1_) / 5
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.cgg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    fork()
    for _ in range(args.steps):
        out = step_resident()
    join()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.cgg_launch_count() - launches0
    if launches == 0:
        # CUDA-graph replay: the host-side counter does not tick; count one eager pass of the same
        # path (identical kernel sequence to the captured one) and scale by the timed steps
        l0 = lib.cgg_launch_count()
        head._runtime(dev)._forward_eager(mf_d, mems_d)
        torch.cuda.synchronize()
        launches = (lib.cgg_launch_count() - l0) * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms = reduce_max(ms, world, dev)
    value = whole_job_value(world, B, args.steps, ms)

    # ---- e2e: public API with HOST buffers; H2D of the step's inputs and D2H of the step's
    # result (last layer's cls / cls_emb / mask logits, what simple_test consumes, head.py:943-945)
    res_h = None
    h2d = mf_h.numel() * mf_h.element_size() + sum(m.numel() * m.element_size() for m in mems_h)

    # Three streams, double-buffered device inputs: the H2D copy of step i+1 and the D2H read of
    # step i-1 overlap the kernels of step i (PCIe is full duplex); every step still pays its own
    # H2D + D2H, and the loop ends with a full synchronise.
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    in_bufs = [(torch.empty_like(mf_d), [torch.empty_like(m) for m in mems_d]) for _ in range(2)]
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    out_free = torch.cuda.Event()

    def upload(slot):
        with torch.cuda.stream(s_in):
            s_in.wait_event(in_free[slot])
            in_bufs[slot][0].copy_(mf_h, non_blocking=True)
            for d, hsrc in zip(in_bufs[slot][1], mems_h):
                d.copy_(hsrc, non_blocking=True)
            in_ready[slot].record(s_in)

    def run_e2e(n):
        nonlocal res_h
        for e in in_free:
            e.record(main)
        out_free.record(main)
        upload(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                upload(slot ^ 1)
            main.wait_event(in_ready[slot])
            cls, emb, mask = head.decoder_forward(in_bufs[slot][0], in_bufs[slot][1])
            in_free[slot].record(main)
            done = torch.cuda.Event()
            done.record(main)
            if res_h is None:
                res_h = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in (cls[-1], emb[-1], mask[-1])]
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                for dst, src in zip(res_h, (cls[-1], emb[-1], mask[-1])):
                    src.record_stream(s_out)
                    dst.copy_(src, non_blocking=True)
                out_free.record(s_out)
        torch.cuda.synchronize()

    run_e2e(2)
    barrier()
    e2e_steps = max(4, args.steps)
    t0 = time.perf_counter()
    run_e2e(e2e_steps)
    barrier()
    e2e_val = whole_job_value(world, B, e2e_steps, 1e3 * reduce_max(time.perf_counter() - t0, world, dev))
    d2h = sum(x.numel() * x.element_size() for x in res_h)

    # ---- roofline of the dominant kernel (the mask einsum, 55% of the path's FLOPs): CUDA events
    # on the launching stream around that stage alone, same inputs, averaged over launches
    rt = head._runtime(dev)
    peaks = measured_peaks()
    reps = 10
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.precision == 'bf16':
        # K2 alone: ONE launch of the tcgen05 GEMM computes the mask logits of all 10 head calls
        # (the mask embeddings of the last forward are still in the workspace); 2.1 GB of output
        # per launch, far larger than L2.
        mask_out = torch.empty((LAYERS + 1, B, Q, H // 4, W // 4), dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            rt.mask_einsum(mf_d, mask_out)
        torch.cuda.synchronize()
        k0.record()
        for _ in range(reps):
            rt.mask_einsum(mf_d, mask_out)
        k1.record()
        torch.cuda.synchronize()
        kern_ms = k0.elapsed_time(k1) / reps
        flops = einsum_flops_per_launch(Q, B) * (LAYERS + 1)
        kname = 'tc_einsum_t_kernel (cta_group::2, queries on TMEM lanes): mask einsum of all 10 head calls, one launch (cgg_mask_einsum)'
        del mask_out
    else:
        x0 = torch.randn((B, Q, C), device=dev)
        for _ in range(3):
            rt.head_call(x0, mf_d, 0, want_bits=False)
        torch.cuda.synchronize()
        k0.record()
        for _ in range(reps):
            rt.head_call(x0, mf_d, 0, want_bits=False)
        k1.record()
        torch.cuda.synchronize()
        kern_ms = k0.elapsed_time(k1) / reps
        flops = einsum_flops_per_launch(Q, B)
        kname = 'cgg_head_call stage (fp32 SIMT heads + mask einsum of one head call)'
    ach = flops / (kern_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'einsum_dram_traffic.json')
    if args.precision == 'bf16' and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get('batch') == B and tj.get('queries') == Q:
            traffic = tj.get('dram_bytes_per_launch')
    roofline = dict(bound='tensor', achieved=ach, peak=peaks['tflops'], unit='TFLOP/s', frac=ach / peaks['tflops'],
                    traffic=traffic, kernel=kname, kernel_ms=kern_ms,
                    algorithmic_bytes=(B * C * (H // 4) * (W // 4) * 2 + (LAYERS + 1) * B * Q * (H // 4) * (W // 4) * 2)
                    if args.precision == 'bf16' else None,
                    peak_source=peaks['source'] + ' (sustained bf16)',
                    whole_path=dict(achieved=flops_per_image(Q) * value / 1e12, unit='TFLOP/s',
                                    frac=flops_per_image(Q) * value / 1e12 / peaks['tflops']))

    if rank == 0:
        cpu = cpu_baseline(Q, os.cpu_count() or 1) if (world == 1 and not args.no_cpu_baseline) else None
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='bf16' if args.precision == 'bf16' else 'f32', data='synthetic', impl='b200',
                    config=dict(workload='configs[1]: COCO-OVIS instance decoder head, Q=%d, 9 layers, 256-d, 8 heads, '
                                         '1024x1024, batch %d per GPU' % (Q, B),
                                batch_per_gpu=B, global_batch=B * world, precision=args.precision,
                                cuda_graph=not args.no_graph, in_flight_batches=n_fly,
                                final_mask_only=bool(args.final_mask_only),
                                single_batch_latency_ms=latency_ms,
                                l2='inputs (%.0f MB per step) larger than L2, no explicit flush' % (h2d / 1e6),
                                flops_per_image=flops_per_image(Q)),
                    e2e=dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=e2e_steps),
                    gpu_launches=int(launches), roofline=roofline, clocks=clocks)
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=16, help='images per GPU per step')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--queries', type=int, default=100)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the path eagerly instead of replaying a CUDA graph')
    ap.add_argument('--final-mask-only', action='store_true',
                    help='opt-in inference shortcut: produce the last head call mask only (NOT the headline contract)')
    ap.add_argument('--in-flight', type=int, default=2,
                    help='batches in flight in the device-resident throughput leg (one head + stream each)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
