#!/usr/bin/env python
"""Headline benchmark: flow.log_prob samples/s of an 8-layer spline-coupling flow, d=64
(BASELINE.json metric / configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--kind quadratic|cubic] [--hidden 64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the CPU arm (oracle port of the reference)

One step = one log_prob pass over the global batch of 2^22 rows (sharded by rows over the N
ranks, no data-path collective: strong scaling, total work fixed as BASELINE.json words it).
`value` is timed with CUDA events with the shard already resident in HBM; `e2e` runs the same
pass from pinned HOST memory through the public API (H2D copy of the rows and D2H copy of the
log-probabilities inside the timed region).  Prints ONE JSON line on rank 0.
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
for p in (ROOT, os.path.join(ROOT, 'tests', 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch

D, LAYERS, BINS = 64, 8, 16
GLOBAL_BATCH = 1 << 22
LOWER, UPPER = -4.0, 4.0
MASKS = ('ordered_right_half', 'ordered_left_half')
METRIC = 'log_prob samples/s, 8-layer spline coupling d=64'
L2_BYTES = 126 << 20


# stdout carries the ONE JSON line and nothing else: the process's fd 1 is pointed at stderr for the whole
# run (NCCL's version banner / NCCL_DEBUG output, any library print) and the result goes to the saved fd.
_RESULT_OUT = None


def _claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(obj):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(obj) + '\n')
    out.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--kind', default='quadratic', choices=['quadratic', 'cubic'])
    ap.add_argument('--hidden', type=int, nargs='+', default=[64])
    ap.add_argument('--batch', type=int, default=GLOBAL_BATCH, help='global batch (rows)')
    ap.add_argument('--cpu-sample', type=int, default=1 << 15, help='rows of the cpu_baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--generic', action='store_true', help='force the CUDA-core path (no tcgen05)')
    ap.add_argument('--micro-rows', type=int, default=1 << 18, help='train: rows per micro-batch')
    ap.add_argument('--tf32', action='store_true', help='train: let the library GEMMs of the conditioner gradients use TF32')
    ap.add_argument('--hybrid', action='store_true', help='train: conditioner through autograd around the element-wise kernels')
    ap.add_argument('--workload', default='logprob', choices=['logprob', 'affine', 'neural', 'train'],
                    help='logprob = BASELINE.json configs[2] (the headline, default); affine = configs[1]; '
                         'neural = configs[3]; train = configs[4] (side measurements, same JSON schema)')
    return ap.parse_args()


def build_layers(kind, hidden):
    """Random-init conditioners with the reference's defaults (nn.Linear init, last bias 0,
    net/mlp.py:46-53), seeded on the CPU so every rank and the CPU baseline share the weights."""
    import stribor_b200 as st
    torch.manual_seed(123)
    P = 3 * BINS - 1 if kind == 'quadratic' else 2 * BINS + 2
    layers = []
    for i in range(LAYERS):
        net = st.net.MLP(D, list(hidden), D * P)
        tr = st.Spline(D, BINS, latent_net=net, lower=LOWER, upper=UPPER, spline_type=kind)
        layers.append(st.Coupling(tr, mask=MASKS[i % 2]))
    return layers


def workload_name(kind, hidden, batch):
    return (f'{kind} st.Spline coupling flow d={D}, {LAYERS} layers alternating {MASKS[0]}/{MASKS[1]}, '
            f'{BINS} bins, MLP{list(hidden)} tanh, box [{LOWER:g},{UPPER:g}], log_prob, global batch {batch}')


# -------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# -------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '50',
                 '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.samples:
            f = [v.strip() for v in line.split(',')]
            if len(f) < 6:
                continue
            try:
                clk, m = float(f[0]), float(f[1])
            except ValueError:
                continue
            mx = m
            if t0 <= ts <= t1 + 0.2:
                sm.append(clk)
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# -------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, all host threads
# -------------------------------------------------------------------------------------------
def cpu_log_prob_rate(spec, kind, rows, chunk=1 << 13, repeats=1):
    from oracle import coupling_flow_oracle as O
    n_threads = len(os.sched_getaffinity(0))
    torch.set_num_threads(n_threads)
    g = torch.Generator().manual_seed(0)
    y = torch.randn(rows, D, generator=g)
    best = None
    with torch.no_grad():
        O.flow_log_prob(spec, y[:256])                       # warm-up
        for _ in range(repeats):
            t = time.perf_counter()
            for i in range(0, rows, chunk):
                O.flow_log_prob(spec, y[i:i + chunk])
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
    return rows / best, n_threads, best


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except Exception:
        pass
    return 'unknown'


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from stribor_b200.spec import spec_from_layers
    spec = spec_from_layers(build_layers(args.kind, args.hidden))
    rows = 1 << 13                                            # bounded sample per step
    from oracle import coupling_flow_oracle as O
    n_threads = len(os.sched_getaffinity(0))
    torch.set_num_threads(n_threads)
    y = torch.randn(rows, D, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        for _ in range(args.warmup):
            O.flow_log_prob(spec, y)
        t = time.perf_counter()
        for _ in range(args.steps):
            O.flow_log_prob(spec, y)
        dt = time.perf_counter() - t
    value = rows * args.steps / dt
    sample = (f'{rows} rows per step of the same workload; oracle port of the reference (torch CPU ops, '
              f'no O(N^2) domain re-check), {n_threads} threads, {cpu_model()}')
    emit({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': workload_name(args.kind, args.hidden, args.batch)},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': n_threads, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    })


# -------------------------------------------------------------------------------------------
# side workloads (BASELINE.json configs[1], [3], [4]); the driver only runs the default one
# -------------------------------------------------------------------------------------------
def run_side(args):
    import torch.distributed as dist
    import stribor_b200 as st
    from stribor_b200 import _ops
    from stribor_b200.parallel import DataParallelNLL, init_from_env, shard_rows

    rank, world, local = init_from_env('nccl')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    torch.manual_seed(123)
    if args.workload == 'affine':
        d, rows_g = 64, 1 << 20
        layers = [st.Coupling(st.Affine(d, latent_net=st.net.MLP(d, [256, 256], 2 * d)), mask=MASKS[i % 2])
                  for i in range(LAYERS)]
        flow = st.NormalizingFlow(st.UnitNormal(d), layers).to(dev).requires_grad_(False)
        name = f'affine coupling flow d=64, 8 layers, MLP[256,256], log_prob + inverse, global batch {rows_g}'
        a, b = shard_rows(rows_g, rank, world)
        torch.manual_seed(rank)
        y = torch.randn(b - a, d, device=dev)

        def step():
            with torch.no_grad():
                flow.log_prob(y)
                flow.inverse(y)
        unit, per_step = 'samples/s (log_prob + inverse per sample)', rows_g
    elif args.workload == 'neural':
        d, B, T = 16, 65536, 64
        layers = [st.ContinuousAffineCoupling(st.net.MLP(d + 1, [64], 2 * d), st.net.TimeLinear(2 * d),
                                              ('ordered_0', 'ordered_1')[i % 2]) for i in range(4)]
        flow = st.NeuralFlow(layers).to(dev).requires_grad_(False)
        name = f'NeuralFlow 4x ContinuousAffineCoupling dim 16, MLP[17->64->32], TimeLinear(32), x [{B},{T},{d}], forward'
        a, b = shard_rows(B, rank, world)
        torch.manual_seed(rank)
        x = torch.randn(b - a, T, d, device=dev)
        t = torch.rand(b - a, T, 1, device=dev)

        def step():
            with torch.no_grad():
                flow(x, t=t)
        unit, per_step = 'rows/s', B * T
    else:
        d, rows_g = 128, args.batch
        P = 3 * BINS - 1
        layers = [st.Coupling(st.Spline(d, BINS, latent_net=st.net.MLP(d, list(args.hidden), d * P), lower=LOWER,
                                        upper=UPPER, spline_type='quadratic'), mask=MASKS[i % 2])
                  for i in range(LAYERS)]
        flow = st.NormalizingFlow(st.UnitNormal(d), layers).to(dev)
        name = (f'NLL training step (fwd+bwd, grads all-reduced) quadratic spline coupling d=128, 8 layers, 16 bins, '
                f'MLP{list(args.hidden)}, global batch {rows_g}')
        a, b = shard_rows(rows_g, rank, world)
        torch.manual_seed(rank)
        y = torch.randn(b - a, d, device=dev)
        if args.tf32:
            torch.backends.cuda.matmul.allow_tf32 = True
        if args.hybrid:
            os.environ['STRIBOR_B200_TRAIN_HYBRID'] = '1'
        name += (f', micro-batches of {args.micro_rows} rows, conditioner-gradient GEMMs in '
                 f'{"tf32" if args.tf32 else "fp32"}' + (', hybrid path' if args.hybrid else ', fused backward kernel'))
        dp = DataParallelNLL(flow, micro_rows=args.micro_rows)

        def step():
            dp.step(y, rows_g)
        unit, per_step = 'samples/s', rows_g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    barrier()
    n0 = _ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    tm = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_step = tm.item() / args.steps
    if rank == 0:
        emit({'metric': f'{args.workload} throughput', 'value': per_step / (ms_step * 1e-3), 'unit': unit,
                          'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
                          'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
                          'data': 'synthetic', 'config': {'workload': name},
                          'gpu_launches': int(_ops.launch_count() - n0)})
    if world > 1:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------
def main():
    args = parse()
    _claim_stdout()
    if args.impl == 'reference':
        run_reference(args)
        return
    if args.workload != 'logprob':
        run_side(args)
        return

    import torch.distributed as dist
    import stribor_b200 as st
    from stribor_b200 import _ops
    from stribor_b200.host import HostPipeline

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f'--gpus {args.gpus} needs torchrun with {args.gpus} ranks')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    layers = build_layers(args.kind, args.hidden)
    flow = st.NormalizingFlow(st.UnitNormal(D), layers).to(dev)
    if args.generic:
        os.environ['STRIBOR_B200_FORCE_GENERIC'] = '1'
    for p in flow.parameters():
        p.requires_grad_(False)

    rows = args.batch // world
    torch.manual_seed(rank)
    y = torch.randn(rows, D, device=dev)                      # 2^22 x 64 fp32 = 1 GiB  (> 126 MB L2)

    def step():
        with torch.no_grad():
            return flow.log_prob(y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from the start of the warm-up (identical load) to the end of the timed
    # region, so that short timed regions (strong scaling at 8 GPUs: ~20 ms) still get samples
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    t_load0 = time.perf_counter()
    for _ in range(args.warmup):
        lp = step()
    assert torch.isfinite(lp).all().item(), 'non-finite log_prob'
    barrier()
    n0 = _ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        lp = step()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    launches = _ops.launch_count() - n0
    ms = e0.elapsed_time(e1)
    if t1 - t_load0 < 0.35:                                # keep the GPU under the same load until a sample lands
        while time.perf_counter() - t_load0 < 0.35:
            step()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    clocks = sampler.stop(t_load0, t1)
    clocks['window'] = 'warm-up + timed steps (same workload)'
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = tmax.item() / args.steps
    value = args.batch / (ms_step * 1e-3)

    # ---- end to end: pinned host rows -> device -> log_prob -> host -------------------------
    e2e = None
    if not args.no_e2e:
        pipe = HostPipeline(flow, D, dev, chunk_rows=max(1 << 16, min(1 << 19, rows // 4)))
        y_host = torch.empty(rows, D, pin_memory=True)
        y_host.copy_(y)
        lp_host = torch.empty(rows, 1, pin_memory=True)
        for _ in range(2):
            pipe.log_prob(y_host, lp_host)
        barrier()
        k = max(2, min(args.steps, 5))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(k):
            pipe.log_prob(y_host, lp_host)
        s1.record()
        barrier()
        te = torch.tensor([s0.elapsed_time(s1)], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        assert torch.allclose(lp_host, lp.cpu(), rtol=0, atol=0), 'e2e result differs from device result'
        e2e = {'value': args.batch / (te.item() / k * 1e-3), 'unit': 'samples/s',
               'h2d_bytes_per_step': rows * D * 4 * world, 'd2h_bytes_per_step': rows * 4 * world,
               'how': f'pinned host rows -> H2D -> flow.log_prob -> D2H, chunks of {pipe.chunk_rows} rows '
                      'double-buffered on two streams'}

    # ---- roofline of the dominant kernel (one fused layer launch) ---------------------------
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        hbm_peak, peak_src = float(pk['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        tf_peak = float(pk.get('bf16_tflops_sustained', pk.get('bf16_tflops', 1394.6)))
    else:
        hbm_peak, peak_src, tf_peak = 6650.0, 'fallback (B200_PROFILING.md)', 1400.0
    launch_ms = ms / max(launches, 1)                         # rank-local average launch duration
    # layer applications one launch performs: 1 (a launch per layer) or all L (whole flow in one launch,
    # the tile stays in shared memory between layers -- DRAM traffic is then ~L times below the algorithmic
    # figure, which by SURVEY 8d stays L * (8d + 8) B per sample)
    layers_per_launch = LAYERS * args.steps / max(launches, 1)
    bytes_per_launch = rows * (8 * D + 8) * layers_per_launch  # per layer: read y + write x + read/write ldj
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    P = 3 * BINS - 1 if args.kind == 'quadratic' else 2 * BINS + 2
    h = [D // 2] + list(args.hidden)
    flops_row = 2 * (sum(a * b for a, b in zip(h[:-1], h[1:])) + h[-1] * (D // 2) * P)
    tflops = rows * flops_row * layers_per_launch / (launch_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
    if os.path.exists(tpath) and args.kind == 'quadratic' and not args.generic:
        tj = json.load(open(tpath))                             # dram__bytes_{read,write}.sum from ncu --set full
        per_row = tj.get('dram_bytes_per_row_chain', tj['dram_bytes_per_row'] * layers_per_launch) if layers_per_launch > 1 \
            else tj['dram_bytes_per_row']
        traffic = per_row * rows
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': achieved / hbm_peak, 'traffic': traffic, 'peak_source': peak_src,
                'kernel': 'fused coupling layer' + (f', {layers_per_launch:g} layers per launch (tile resident in shared memory)'
                                                    if layers_per_launch > 1 else ' (one launch per layer)'),
                'layers_per_launch': layers_per_launch,
                'bytes_per_launch': bytes_per_launch, 'launch_ms': launch_ms,
                'note': 'BASELINE metric asks for % of HBM peak on L*(8d+8) B/sample; the kernel is '
                        'bound by the conditioner contraction + spline epilogue, see tensor figure'}
    roofline_tensor = {'bound': 'tensor', 'achieved': tflops, 'peak': tf_peak, 'unit': 'TFLOP/s',
                       'frac': tflops / tf_peak,
                       'note': 'masked-minimum conditioner flops (SURVEY 8d); peak = measured bf16 dense'}

    out = {
        'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.kind, args.hidden, args.batch),
                   'rows_per_gpu': rows, 'l2_policy': 'inputs larger than L2 (1 GiB of rows per pass)',
                   'path': 'generic' if args.generic else 'auto'},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
        'roofline': roofline, 'roofline_tensor': roofline_tensor,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from stribor_b200.spec import spec_from_layers
        spec = spec_from_layers([l.cpu() for l in build_layers(args.kind, args.hidden)])
        v, cores, secs = cpu_log_prob_rate(spec, args.kind, args.cpu_sample)
        out['cpu_baseline'] = {
            'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
            'sample': f'{args.cpu_sample} rows of the same workload in chunks of 8192 ({secs:.1f} s); oracle '
                      f'port of the reference on torch CPU ops, {cores} threads, {cpu_model()}; the '
                      'unmodified reference has an O(N^2) domain re-check on the quadratic path '
                      '(57 samples/s at 256-row chunks, BASELINE.md) that the port omits'}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
