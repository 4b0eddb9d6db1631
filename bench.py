#!/usr/bin/env python
"""Benchmark of the InvertAvatar generator-forward hot path (BASELINE.json metric: 512^2 avatar frames/s at
128^2 neural render x 48+48 depth samples).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)
    python bench.py --workload c3 ...                        # BASELINE configs[2]: eval_seq.py few-shot path (encoder)

Workload c2 (default, BASELINE configs[1]): a "step" renders one batch of `--batch` (default 8) frames per GPU: mapping
(z,c -> ws) + TriPlaneGenerator.synthesis (3 StyleGAN2 backbones, UV rasterize/stitch, fused volume renderer,
super-resolution) on synthetic latents / cameras / UV mesh conditions and random-init weights of the reference
architecture.  Frames are independent, so N GPUs run N x batch frames (weak scaling); the only collective is the
all-gather of the final images.

Workload c3 (BASELINE configs[2], eval_seq.py:164-212): a "step" is one identity: e4e encode (B=1) + the two backbones +
inversionNet.AR_eval_forward over T=4 reference frames (ConvGRU/UNet encoders, a 4-frame 128^2 x 48+48 render inside) +
4 x synthesis_withTexture (the per-frame driver) -> 4 frames of 512^2.  The ConvGRU state is sequential in T and
train-mode BatchNorm couples the T frames, so identities are the unit that shards: N GPUs = N replicas.

One JSON line on stdout (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same through the public
API with HOST inputs (pinned H2D every step) and a D2H copy of the frames every step (uint8 HWC as the reference
scripts write them, reenact_avatar_next3d.py:117-131; --e2e-f32 returns the fp32 images instead).  `parity` = max-abs /
PSNR of GPU frames against the CPU oracle frames the cpu_baseline leg renders from the same inputs.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = '512x512 avatar frames/sec (128x128 neural render x 48+48 depth samples)'
METRIC_C3 = '512x512 avatar frames/sec, eval_seq few-shot path (e4e encode + AR_eval_forward T=4 + 4 x synthesis_withTexture per identity)'
UNIT = 'frames/s'
T_C3 = 4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=['c2', 'c3'])
    ap.add_argument('--batch', type=int, default=8, help='frames per GPU per step (workload c2)')
    ap.add_argument('--res', type=int, default=128, help='neural rendering resolution')
    ap.add_argument('--depth', type=int, default=48, help='coarse = importance depth samples per ray')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--c3-inflight', type=int, default=1, help='workload c3: identities in flight per GPU and step, each replaying its own graph on its own stream (throughput mode; default 1 = one identity at a time)')
    ap.add_argument('--c3-per-frame', action='store_true', help='workload c3: render the T driven frames with T batch-1 synthesis_withTexture calls (the loop of eval_seq.py:212) instead of one batch-T call')
    ap.add_argument('--no-graph', action='store_true', help='issue the step eagerly (Python + ctypes launches) instead of replaying its CUDA graph')
    ap.add_argument('--e2e-f32', action='store_true', help='end-to-end leg reads back the fp32 images instead of uint8 HWC frames')
    return ap.parse_args()


def _conv_precision():
    if os.environ.get('IA_CONV_PRECISION', 'auto') == 'bf16x3':
        return 'bf16x3 split operands (hi*hi+hi*lo+lo*hi) in every convolution, fp32 accumulate in TMEM'
    return ('per layer (profiles/r2_conv_precision_probe*.json): backbone 3x3 convolutions single-pass fp16 operands, super-resolution and '
            'ToRGB convolutions bf16x3 split operands (hi*hi+hi*lo+lo*hi); fp32 accumulate in TMEM; encoder convolutions bf16x3')


def workload_config(args, n_gpus, frames_per_step=None):
    fps = args.batch if frames_per_step is None else frames_per_step
    if args.workload == 'c3':
        return {'workload': f'eval_seq.py few-shot path (BASELINE configs[2]): encode B=1 + AR_eval_forward T={T_C3} ({args.res}^2 x {args.depth}+{args.depth} render inside) '
                            + (f'+ {T_C3} x synthesis_withTexture (batch 1 each)' if getattr(args, 'c3_per_frame', False) else f'+ one batch-{T_C3} synthesis_withTexture call for the {T_C3} driven frames')
                            + ', random-init inversionNet + generator, synthetic images/UV/cameras',
                'frames_per_gpu_per_step': T_C3 * getattr(args, 'c3_inflight', 1), 'global_frames_per_step': T_C3 * n_gpus * getattr(args, 'c3_inflight', 1),
                'identities_per_gpu_per_step': getattr(args, 'c3_inflight', 1),
                'neural_res': args.res, 'depth_samples': [args.depth, args.depth],
                'parallelism': f'{n_gpus} replica(s): identities are independent, the ConvGRU state is sequential inside one',
                'conv_precision': _conv_precision(), 'gather': 'none',
                'l2': 'working set per step exceeds the 126 MB L2; no explicit flush'}
    return {'workload': f'Next3D++ reenactment 512^2, {args.res}^2 neural x {args.depth}+{args.depth} depth, batch {fps}/GPU '
                        f'(BASELINE configs[1]), random-init generator, synthetic latents/cameras/UV',
            'frames_per_gpu_per_step': fps, 'global_frames_per_step': fps * n_gpus,
            'neural_res': args.res, 'depth_samples': [args.depth, args.depth], 'parallelism': f'dp{n_gpus} (frames sharded, weights replicated)',
            'conv_precision': _conv_precision(), 'gather': getattr(args, 'gather_kind', 'none'),
            'l2': 'working set per step (>10 GB of activations) exceeds the 126 MB L2; no explicit flush'}


# ----------------------------------------------------------------------------------------------------------------
# algorithmic work: 2*MAC, transposed (up=2) convolutions counted at input resolution
# ----------------------------------------------------------------------------------------------------------------
def conv_flops_per_frame(G):
    from invertavatar_b200 import stylegan2 as sg
    total = 0
    for m in G.modules():
        if isinstance(m, sg.SynthesisLayer):
            r_in = m.resolution // m.up
            total += 2 * r_in * r_in * m.in_channels * m.out_channels * 9
    for name, blk in G.named_modules():
        if isinstance(blk, sg.SynthesisBlock) and hasattr(blk, 'torgb'):
            r = blk.resolution
            total += 2 * r * r * blk.torgb.in_channels * blk.torgb.out_channels
    return total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200', '-i', str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, col in (('hw_slowdown', 5), ('hw_thermal_slowdown', 6), ('sw_thermal_slowdown', 7), ('sw_power_cap', 8)):
                if len(r) > col and r[col].lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[3]))
            except (ValueError, IndexError):
                pass
        return {'sm_mhz': (sm[len(sm) // 2] if sm else None), 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm),
                'sm_mhz_min': (sm[0] if sm else None), 'power_w_max': (max(pw) if pw else None)}


def _psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float('inf') if mse == 0 else 10.0 * math.log10(4.0 / mse)      # images span [-1, 1]: peak-to-peak 2


def _parity(got, ref, what):
    err = float((got.float() - ref.float()).abs().max())
    return {'max_abs': err, 'psnr_db': _psnr(got.float(), ref.float()), 'frames_compared': int(got.shape[0]),
            'tolerance': '1e-3 max-abs, PSNR > 50 dB (BASELINE north_star)', 'ok': bool(err <= 1e-3 and _psnr(got.float(), ref.float()) > 50.0),
            'config': what}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


def _roofline(conv_ms, conv_launches, flops, steps, extra):
    """Dominant kernel = the tcgen05 convolution: algorithmic FLOPs of the timed steps / its summed launch durations (CUDA
    events on the launching stream around every launch, ia_profile_begin/report)."""
    peaks = _peaks()
    peak = float(peaks.get('bf16_tflops_sustained', 1400.0))
    burst = float(peaks.get('bf16_tflops', 1650.0))
    achieved = flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    r = {'kernel': 'conv_tc2_kernel / conv_tc_kernel (tcgen05 implicit-GEMM modulated convolution)', 'bound': 'tensor', 'achieved': achieved,
         'peak': peak, 'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peaks else 'fallback 1.4 PFLOP/s sustained (of fallback)',
         'unit': 'TFLOP/s', 'frac': achieved / peak, 'peak_burst': burst, 'frac_burst': achieved / burst,
         'launches_per_step': conv_launches / steps, 'avg_launch_ms': conv_ms / max(1, conv_launches)}
    r.update(extra)
    return r


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated in oracle/ (the reference is Python/torch and is not present on the GPU
# box; the oracle is pinned to it by tests/golden).  Used as cpu_baseline of the B200 arm and as --impl reference.
# ----------------------------------------------------------------------------------------------------------------
def cpu_c2(args, frames, repeats, warmup=1):
    """-> (frames/s, cores, per-step seconds, image [frames,3,512,512] of the last step)."""
    import torch
    from invertavatar_b200 import synth
    from invertavatar_b200.triplane import TriPlaneGenerator
    from oracle import triplane as o_tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    G = TriPlaneGenerator(**synth.generator_kwargs(args.depth, args.depth)).eval().requires_grad_(False)
    synth.randomize_noise_and_wavg(G)
    sd = G.state_dict()
    kw = G.rendering_kwargs
    z, cond, c, uv = synth.latents(frames), synth.frontal_camera(frames), synth.cameras(frames), synth.uvcoords_image(frames)
    jit = synth.depth_jitter(frames, args.res * args.res, args.depth)

    def step():
        ws = o_tp.mapping(sd, z, cond, kw, truncation_psi=0.7, truncation_cutoff=14)
        return o_tp.synthesis(sd, ws, c, uv, kw, jit, evaluation=True, neural_rendering_resolution=args.res)['image']
    img = None
    with torch.no_grad():
        for _ in range(warmup):
            step()
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            img = step()
            times.append(time.perf_counter() - t0)
    return frames * len(times) / sum(times), cores, times, img


def build_c3(args, device=None):
    """The inversionNet of eval_seq.py:83-97 (random-init, reference architecture and mode flags) and the synthetic
    T-frame identity of SURVEY 8(d)."""
    import torch
    from invertavatar_b200 import synth
    from invertavatar_b200.encoder import inversionNet
    from invertavatar_b200.triplane import TriPlaneGenerator
    torch.manual_seed(0)
    G = TriPlaneGenerator(**synth.generator_kwargs(args.depth, args.depth)).eval().requires_grad_(False)
    synth.randomize_noise_and_wavg(G)
    torch.manual_seed(1)
    net = inversionNet(generator=G, encoding_triplane=True, encoding_texture=True).train().requires_grad_(False)
    synth.randomize_encoder(net)
    for u in (net.unet_encoder.triplane_unet, net.unet_encoder.texture_unet):
        u.input_layer.eval(); u.body.eval()
    net.generator.neural_rendering_resolution = args.res
    x, c, v = synth.encoder_inputs(T_C3)
    if device is not None:
        net = net.to(device)
    return net, x, c, v


def c3_draws(args):
    """The pinned random draws of one c3 step: (jitter, u) of the T-frame render inside AR_eval_forward (evaluation=False)
    and the jitter of each of the T driven frames (evaluation=True)."""
    from invertavatar_b200 import synth
    rays = args.res * args.res
    return (synth.depth_jitter(T_C3, rays, args.depth, seed=20), synth.importance_u(T_C3, rays, args.depth, seed=30),
            [synth.depth_jitter(1, rays, args.depth, seed=40 + i) for i in range(T_C3)])


def cpu_c3(args, repeats, warmup=0):
    """One identity through the oracle port -> (frames/s, cores, per-step seconds, the T driven frames [T,3,512,512])."""
    import torch
    from oracle import encoder as o_enc
    from oracle import stylegan2 as o_sg
    from oracle import triplane as o_tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net, x, c, v = build_c3(args)
    sd = {k: t.clone() for k, t in net.state_dict().items()}      # (the oracle's train-mode BatchNorm does not touch running stats)
    kw = net.generator.rendering_kwargs
    gsd = o_sg.sub(sd, 'generator')
    jit_ar, u_ar, jit_frames = c3_draws(args)
    uvimg = v['uvcoords_image']

    def step():
        ws = o_enc.encode(sd, x['image'][:1], training=True)
        tex = o_sg.synthesis_network(o_sg.sub(gsd, 'texture_backbone.synthesis'), ws, return_list=True)
        sta = o_sg.synthesis_network(o_sg.sub(gsd, 'backbone.synthesis'), ws, return_list=True)
        upd, _ = o_enc.ar_eval_forward(sd, x, c, uvimg, ws, [None, None], kw, jit_ar, u_ar, e4e_results={'w': ws, 'texture': tex, 'static': sta},
                                       neural_rendering_resolution=args.res)
        frames = [o_tp.synthesis_with_texture(gsd, ws, upd['texture'], c[i:i + 1], uvimg[i:i + 1], kw, jit_frames[i], static_feats=upd['static'],
                                              evaluation=True, neural_rendering_resolution=args.res)['image'] for i in range(T_C3)]
        return torch.cat(frames)
    img = None
    with torch.no_grad():
        for _ in range(warmup):
            step()
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            img = step()
            times.append(time.perf_counter() - t0)
    return T_C3 * len(times) / sum(times), cores, times, img


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    warm = 1 if args.warmup > 0 else 0
    if args.workload == 'c3':
        steps = max(1, min(args.steps, 2))
        fps, cores, times, _ = cpu_c3(args, steps, warmup=0)
        warm = 0
        sample = f'{steps} identity step(s) of the same workload ({T_C3} frames each), no warm-up (reference algorithm = oracle port, torch CPU fp32, {cores} threads)'
        cfg, metric = workload_config(args, args.gpus), METRIC_C3
    else:
        # the batch the B200 arm runs (default 8 frames per step), bounded to 3 timed steps so that the run ends within minutes
        steps = max(1, min(args.steps, 3))
        fps, cores, times, _ = cpu_c2(args, args.batch, steps, warmup=warm)
        sample = f'{args.batch} frames/step x {steps} steps of the same workload after {warm} warm-up (reference algorithm = oracle port, torch CPU fp32, {cores} threads)'
        cfg, metric = workload_config(args, args.gpus), METRIC
        cfg['parallelism'] = f'host CPU, {cores} threads (rank 0 only)'
    line = {'impl': 'reference', 'metric': metric, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps, 'warmup': warm,
            'ms_per_step': 1000.0 * sum(times) / len(times), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    _emit_line(line)


# ----------------------------------------------------------------------------------------------------------------
class _Harness:
    """Timing plumbing shared by the two workloads: barrier + synchronize on both sides, CUDA events, max over ranks."""

    def __init__(self, dev, world, copy_stream):
        self.dev, self.world, self.copy_stream = dev, world, copy_stream

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        import torch
        import torch.distributed as dist
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        # Host flow control: at most `depth` steps are in flight.  Without it the host (8 ms of launch work per 17 ms step) runs
        # up to a launch queue (~8 steps) ahead of the device, every step in flight pins its own set of cross-stream activation
        # blocks in torch's caching allocator (blocks handed between the backbone streams are reusable only after their
        # record_stream events complete), and the pool grows by cudaMalloc in the middle of a timed region -- measured as a
        # one-off 100-200 ms stall in one region out of five.  The device still runs back to back (two steps are queued).
        depth, marks = 2, []
        for _ in range(steps):
            if len(marks) >= depth:
                marks[len(marks) - depth].synchronize()
            fn()
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append(ev)
        torch.cuda.current_stream().wait_stream(self.copy_stream)   # the last read-back belongs to the timed region
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        prev, per_step = e0, []
        for ev in marks:                       # per-step device time (diagnostic: a stall shows up as one long step)
            per_step.append(prev.elapsed_time(ev))
            prev = ev
        self.last_step_ms = per_step
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


class _Readback:
    """Double-buffered device->host read-back on a second stream: the copy of step i overlaps the kernels of step i+1; a host
    buffer is reused only after its previous copy has landed."""

    def __init__(self, like, copy_stream):
        import torch
        self.copy_stream = copy_stream
        self.bufs = [torch.empty(like.shape, dtype=like.dtype).pin_memory() for _ in range(2)]
        self.done = [None, None]
        self.i = 0
        self.bytes = like.numel() * like.element_size()

    def push(self, t):
        import torch
        k = self.i & 1
        self.i += 1
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            if self.done[k] is not None:
                self.done[k].synchronize()
            self.bufs[k].copy_(t, non_blocking=True)
            t.record_stream(self.copy_stream)
            d = torch.cuda.Event()
            d.record(self.copy_stream)
            self.done[k] = d
        return self.bufs[k]


def run_b200(args):
    import torch
    import torch.distributed as dist
    from invertavatar_b200 import _C

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback for the product path)'
    _C.lib()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    try:
        line = run_c3(args, rank, world, local, dev) if args.workload == 'c3' else run_c2(args, rank, world, local, dev)
    finally:
        if world > 1:
            dist.destroy_process_group()
    if rank == 0:
        _emit_line(line)


def run_c2(args, rank, world, local, dev):
    import torch
    import torch.distributed as dist
    from invertavatar_b200 import synth
    from invertavatar_b200 import runtime as rt
    from invertavatar_b200.triplane import TriPlaneGenerator

    B, res, D = args.batch, args.res, args.depth
    torch.manual_seed(0)
    G = TriPlaneGenerator(**synth.generator_kwargs(D, D)).eval().requires_grad_(False)
    synth.randomize_noise_and_wavg(G)
    G = G.to(dev)
    first = rank * B     # this rank's frames of the global batch
    z_h = synth.latents(B, first).pin_memory()
    cond_h = synth.frontal_camera(B).pin_memory()
    c_h = synth.cameras(B, first).pin_memory()
    uv_h = synth.uvcoords_image(B, first).pin_memory()
    z, cond, c, uv = z_h.to(dev), cond_h.to(dev), c_h.to(dev), uv_h.to(dev)
    gathered = torch.empty((world * B, 3, 512, 512), dtype=torch.float32, device=dev) if world > 1 else None
    # The one exchange step of the path: gather the final images.  Preferred: the last ToRGB kernel stores its
    # frames straight into every rank's gathered buffer over NVLink (symmetric memory, NVSwitch multicast when available) and a
    # cross-rank barrier closes the step; fallback (IA_GATHER=nccl or no symmetric memory): one NCCL all-gather.
    peer = None
    gather_kind = 'none'
    if world > 1:
        gather_kind = 'nccl all_gather_into_tensor'
        if os.environ.get('IA_GATHER', 'p2p') != 'nccl':
            try:
                from invertavatar_b200.parallel import PeerFrameGather
                peer = PeerFrameGather(B, (3, 512, 512), device=dev)
                gather_kind = 'fused into the last ToRGB kernel: ' + ('multimem.st over the NVSwitch multicast mapping' if peer.mc_ptr else
                                                                      'stores to peer-mapped symmetric memory') + ' + barrier (double-buffered slots)'
            except Exception as ex:   # symmetric memory unavailable on this box / build
                peer = None
                gather_kind += f' (symmetric memory unavailable: {type(ex).__name__})'

    def frame_batch(z, cond, c, uv, jitter=None):
        ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
        kw = dict(neural_rendering_resolution=res, noise_mode='const', evaluation=True)
        if jitter is not None:
            kw['depth_jitter'] = jitter
        if peer is not None:
            with peer.sink():
                img = G.synthesis(ws, c, {'uvcoords_image': uv}, **kw)['image']
            peer.barrier()
            return img
        img = G.synthesis(ws, c, {'uvcoords_image': uv}, **kw)['image']
        if world > 1:
            dist.all_gather_into_tensor(gathered, img.contiguous())
        return img

    # The step is issued as ONE CUDA graph replay (invertavatar_b200.graphs.GraphedCall over mapping + synthesis: the same 105
    # kernels, captured once): issuing them one by one through Python + ctypes costs about as much host time as the step takes on
    # the device, which is what made the end-to-end number fall behind the device number when 8 ranks share one host.  The
    # cross-rank part of the step (peer barrier / NCCL gather) stays outside the graph; with the fused peer gather there is one
    # graph per slot of the double-buffered gathered tensor.  --no-graph issues the launches eagerly.
    graphs = None
    launches_per_step = None
    graph_error = None
    if not args.no_graph:
        from invertavatar_b200.graphs import GraphedCall

        def compute(z, cond, c, uv):
            ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
            kw = dict(neural_rendering_resolution=res, noise_mode='const', evaluation=True)
            if peer is not None:
                with peer.sink():
                    return G.synthesis(ws, c, {'uvcoords_image': uv}, **kw)['image']
            return G.synthesis(ws, c, {'uvcoords_image': uv}, **kw)['image']
        with torch.no_grad():
            compute(z, cond, c, uv)
            torch.cuda.synchronize()
            rt.reset_launch_count()
            compute(z, cond, c, uv)
            launches_per_step = rt.launch_count()
            try:
                graphs = []
                for k in range(2 if peer is not None else 1):
                    if peer is not None:
                        peer.cur = k
                    graphs.append(GraphedCall(compute, dict(z=z, cond=cond, c=c, uv=uv)))
            except Exception as ex:      # capture refused on this box / driver: issue eagerly and say so in the line
                graphs, graph_error = None, f'{type(ex).__name__}: {ex}'[:200]
                torch.cuda.synchronize()
            if peer is not None:
                peer.cur, peer.last = 0, 0
        if world > 1:      # every rank must take the same path (the peer barrier / NCCL gather are matched calls)
            flag = torch.tensor([0 if graphs is None else 1], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                graphs = None

    def frame_batch_graph(new_inputs=None):
        g = graphs[peer.cur if peer is not None else 0]
        img = g(**new_inputs) if new_inputs else g()
        if peer is not None:
            peer.barrier()
        elif world > 1:
            dist.all_gather_into_tensor(gathered, img.contiguous())
        return img

    def step_resident():
        if graphs is not None:
            return frame_batch_graph()
        return frame_batch(z, cond, c, uv)

    def step_resident_eager():
        return frame_batch(z, cond, c, uv)

    # End-to-end step: pinned-host inputs are uploaded on a copy stream one step ahead (the upload of step i+1 overlaps the
    # kernels of step i; the compute stream waits on the upload's event), the frames go back to pinned host memory on the
    # same copy stream (double-buffered).  Every copy is enqueued inside the timed region and the closing synchronize waits
    # for all of them.  The frames are read back as the reference scripts write them: uint8 HWC (layout_grid with
    # float_to_uint8, reenact_avatar_next3d.py:117-131) -- 0.79 MB per frame instead of 3.1 MB of fp32.
    copy_stream = torch.cuda.Stream(device=dev)
    h = _Harness(dev, world, copy_stream)
    host_in = (z_h, cond_h, c_h, uv_h)
    stage = {'next': None}

    def upload():
        with torch.cuda.stream(copy_stream):
            ts = tuple(t.to(dev, non_blocking=True) for t in host_in)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ts, ev

    rb = {'o': None}

    def step_e2e():
        cur = torch.cuda.current_stream()
        if stage['next'] is None:
            stage['next'] = upload()
        (zz, cc, c2, uu), ev = stage['next']
        cur.wait_event(ev)
        for t in (zz, cc, c2, uu):
            t.record_stream(cur)
        stage['next'] = upload()               # next step's inputs travel while this step computes
        img = frame_batch_graph(dict(z=zz, cond=cc, c=c2, uv=uu)) if graphs is not None else frame_batch(zz, cc, c2, uu)
        out = img if args.e2e_f32 else rt.layout_grid_u8(img, grid_w=B, grid_h=1)
        if rb['o'] is None:
            rb['o'] = _Readback(out, copy_stream)
        rb['o'].push(out)
        return img

    parity = None
    with torch.no_grad():
        # nvidia-smi sampler (every 200 ms) is started BEFORE the warm-up: its start-up (driver initialisation of a second
        # process, ~0.5 s) stalls kernel launches for tens of milliseconds and must not land in a timed region
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
            time.sleep(1.0)
        for _ in range(max(args.warmup, 3)):
            step_resident()
        if peer is not None:   # the fused gather must reproduce the NCCL gather bit for bit
            img = step_resident()
            dist.all_gather_into_tensor(gathered, img.contiguous())
            torch.cuda.synchronize()
            ok = torch.tensor([1 if torch.equal(gathered, peer.tensor) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            assert int(ok.item()) == 1, 'fused peer-memory gather differs from the NCCL all-gather'
        clocks.rows.clear()     # keep only the samples taken from here on: both timed regions (resident + end-to-end)
        rt.reset_launch_count()
        ms = h.timed(step_resident, args.steps)
        step_ms = sorted(h.last_step_ms)
        launches = rt.launch_count() if graphs is None else launches_per_step * args.steps
        for _ in range(2):
            step_e2e()
        ms_e2e = h.timed(step_e2e, args.steps)
        step_ms_e2e = sorted(h.last_step_ms)
        clk = clocks.stop() if rank == 0 else None

        roofline = breakdown = None
        if not args.no_roofline:
            # same steps again with every launch bracketed by CUDA events on the launching stream (every rank runs them so
            # that the image all-gather stays matched; rank 0 reports)
            h.barrier()
            from invertavatar_b200 import triplane as _tp
            _tp.set_backbone_streams(False)      # per-kernel durations are taken with one kernel at a time on the device
            rt.flop_count_begin()
            rt.profile_begin()
            for _ in range(args.steps):
                step_resident_eager()       # eager launches: the event brackets of ia_profile_begin live in the host-side launch path
            rep = rt.profile_report()
            counted = rt.flop_count_end()
            _tp.set_backbone_streams(None)
            h.barrier()
            if rank == 0:
                conv = rep.get('ia_conv_tc', {'ms': 0.0, 'launches': 0})
                per_frame = conv_flops_per_frame(G)
                # numerator: the convolutions actually launched (runtime counter) -- the texture backbone's dead 256^2 block and dead
                # skip images are not launched, so they are not counted either (the module-based figure below includes them)
                flops = counted['algorithmic']
                roofline = _roofline(conv['ms'], conv['launches'], flops, args.steps, {
                    # DRAM bytes of one launch from the committed `ncu --set full` capture (profiles/r1_conv_same_full_v7.txt): the
                    # 128->128 @512^2 layer at batch 8 reads 1.085 GB and writes 1.030 GB; its algorithmic bytes are the bf16 hi/lo
                    # operand in (1.074 GB) and the next layer's operand out (1.074 GB) -- no re-reads
                    'traffic': 1.127e9, 'traffic_launch': 'conv_tc2 128->128 @512x512 (SR block1 conv1, ToRGB fused), batch 8: dram read 1.100 GB + write 0.027 GB '
                    '(profiles/r2_conv_full_v16.txt); algorithmic 1.074 GB of bf16 hi/lo operand + 0.6 MB of weights + 25 MB of rgb',
                    'algorithmic_flops_per_frame': counted['algorithmic'] / (B * args.steps),
                    'reference_flops_per_frame_incl_dead_layers': per_frame,
                    'issued_mma_flops_per_frame': counted['issued_mma'] / (B * args.steps),
                    'issued_tflops': counted['issued_mma'] / (conv['ms'] * 1e-3) / 1e12 if conv['ms'] > 0 else 0.0,
                    'note': 'achieved/frac count algorithmic FLOPs (fp32 semantics); issued_* count the tensor-core MMAs really issued '
                            '(split terms x padded channels), so frac <= algorithmic/issued by construction'})
                roofline['issued_frac'] = roofline['issued_tflops'] / roofline['peak']
                tot = sum(v['ms'] for v in rep.values())
                breakdown = {k: {'ms_per_step': v['ms'] / args.steps, 'launches_per_step': v['launches'] / args.steps, 'share': v['ms'] / tot}
                             for k, v in sorted(rep.items(), key=lambda kv: -kv[1]['ms'])}

        cpu = None
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            # CPU leg (bounded sample: 1 frame/step) and parity: frame 0 of one more GPU batch, rendered with the depth jitter the
            # oracle gets, against the oracle's frame
            fps, cores, times, ref_img = cpu_c2(args, 1, 2, warmup=1)
            cpu = {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': f'1 frame/step x 2 steps of the same workload after 1 warm-up (oracle port of the reference, torch CPU fp32, {cores} threads)'}
            jit = synth.depth_jitter(B, res * res, D).to(dev)
            got = frame_batch(z, cond, c, uv, jitter=jit)[:1].float().cpu()
            parity = _parity(got, ref_img, f'frame 0 of the batch-{B} GPU step vs the oracle frame of the cpu_baseline leg (same latent, camera, UV, depth jitter)')

    frames = B * world * args.steps
    value = frames / (ms * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host_in)
    d2h = rb['o'].bytes
    if rank != 0:
        return None
    args.gather_kind = gather_kind
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args, world), 'clocks': clk,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps,
                    'readback': 'fp32 NCHW images' if args.e2e_f32 else 'uint8 HWC frames (layout_grid, reenact_avatar_next3d.py:117-131)'},
            'gpu_launches': launches,
            'issue': ('eager (Python + ctypes launches)' + (f'; graph capture failed: {graph_error}' if not args.no_graph and graph_error else '')
                      if graphs is None else
                      f'one CUDA graph replay per step (invertavatar_b200.graphs.GraphedCall over G.mapping + G.synthesis: {launches_per_step} kernels of '
                      'libinvertavatar_b200.so per step, counted on an eager step and captured once); gpu_launches = kernels per step x steps'),
            'step_ms': {'median': step_ms[len(step_ms) // 2], 'min': step_ms[0], 'max': step_ms[-1],
                        'e2e_median': step_ms_e2e[len(step_ms_e2e) // 2], 'e2e_max': step_ms_e2e[-1]}}
    if roofline is not None:
        line['roofline'] = roofline
        line['kernel_breakdown'] = breakdown
    if cpu is not None:
        line['cpu_baseline'] = cpu
        line['parity'] = parity
        assert parity['ok'], f'GPU frame does not match the oracle: {parity}'
    return line


def run_c3(args, rank, world, local, dev):
    import torch
    from invertavatar_b200 import runtime as rt

    net, x_h, c_h, v_h = build_c3(args, dev)
    G = net.generator
    host_in = {'image': x_h['image'].pin_memory(), 'uv': x_h['uv'].pin_memory(), 'c': c_h.pin_memory(), 'uvimg': v_h['uvcoords_image'].pin_memory()}
    res_in = {k: t.to(dev) for k, t in host_in.items()}

    def identity(inp, draws=None):
        """eval_seq.py:164-212 for one identity: -> the T driven frames [T,3,512,512]."""
        x = {'image': inp['image'], 'uv': inp['uv']}
        c, v = inp['c'], {'uvcoords_image': inp['uvimg']}
        ws = net.encode(x['image'][:1])
        # the two batch-1 backbone passes of eval_seq.py:170-171 are independent chains of small launches: side by side on two streams
        cur = torch.cuda.current_stream(dev)
        side = rt.side_streams(dev, 7)[6]
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            sta = G.backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
        tex = G.texture_backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
        cur.wait_stream(side)
        for t in sta:
            t.record_stream(cur)
        if draws is not None:
            G.renderer.depth_jitter, G.renderer.importance_u = draws[0], draws[1]
        upd, _ = net.AR_eval_forward(x, c, v, ws, [None, None], e4e_results={'w': ws, 'texture': tex, 'static': sta}, return_fake=False)
        if not args.c3_per_frame:
            # the T driven frames are independent given the updated features: one batch-T synthesis_withTexture call (the engine's API
            # takes a batch, as AR_eval_forward's own T-frame render does) instead of the script's per-frame loop (eval_seq.py:212)
            if draws is not None:
                G.renderer.depth_jitter = torch.cat(list(draws[2]), dim=0)
            return G.synthesis_withTexture(ws.expand(T_C3, -1, -1), [f.expand(T_C3, -1, -1, -1) for f in upd['texture']], c, v, noise_mode='const',
                                           static_feats=[f.expand(T_C3, -1, -1, -1) for f in upd['static']], evaluation=True)['image']
        frames = []
        for i in range(T_C3):
            if draws is not None:
                G.renderer.depth_jitter = draws[2][i]
            frames.append(G.synthesis_withTexture(ws, upd['texture'], c[i:i + 1], {'uvcoords_image': v['uvcoords_image'][i:i + 1]}, noise_mode='const',
                                                  static_feats=upd['static'], evaluation=True)['image'])
        return torch.cat(frames)

    # The identity step is ~2 400 launches whose host cost exceeds their device time: the public helper for such sequences,
    # invertavatar_b200.graphs.GraphedCall, captures it once into a CUDA graph and replays it per identity (same kernels, same
    # numbers; --no-graph issues them eagerly).
    graphed = None
    launches_per_step = None
    if not args.no_graph:
        from invertavatar_b200.graphs import GraphedCall
        with torch.no_grad():
            identity(res_in)                     # (first call builds the weight packs)
            rt.reset_launch_count()
            identity(res_in)
            launches_per_step = rt.launch_count()
        graphed = GraphedCall(lambda **kw: identity(kw), res_in)
    # throughput mode (--c3-inflight n > 1): n independent identities per step, each on its own stream with its own graph -- the
    # identity step is a chain of short, latency-bound launches (few CTAs each), so concurrent chains fill the idle SMs
    n_fl = max(1, int(args.c3_inflight)) if graphed is not None else 1
    extra = [GraphedCall(lambda **kw: identity(kw), res_in) for _ in range(n_fl - 1)] if n_fl > 1 else []
    fl_streams = [torch.cuda.Stream(device=dev) for _ in range(n_fl - 1)]

    def replay_all(inputs=None):
        # -> list of the n identities' frame tensors; identity 0 on the caller's stream, the others on their own streams
        cur = torch.cuda.current_stream(dev)
        outs = [None] * n_fl
        for k, (g, s_) in enumerate(zip(extra, fl_streams)):
            s_.wait_stream(cur)
            with torch.cuda.stream(s_):
                outs[k + 1] = g(**inputs) if inputs else g()
        outs[0] = graphed(**inputs) if inputs else graphed()
        for s_ in fl_streams:
            cur.wait_stream(s_)
        return outs

    def step_resident():
        if graphed is None:
            return identity(res_in)
        return replay_all()[0] if n_fl > 1 else graphed()

    copy_stream = torch.cuda.Stream(device=dev)
    h = _Harness(dev, world, copy_stream)
    stage = {'next': None}
    rb = {'o': None}

    def upload():
        with torch.cuda.stream(copy_stream):
            ts = {k: t.to(dev, non_blocking=True) for k, t in host_in.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ts, ev

    def step_e2e():
        cur = torch.cuda.current_stream()
        if graphed is not None and n_fl > 1:
            imgs = replay_all(host_in)           # every identity uploads its own inputs on its own stream, then replays
            img = torch.cat(imgs) if args.e2e_f32 else None
            out = img if args.e2e_f32 else rt.layout_grid_u8(torch.cat(imgs), grid_w=T_C3 * n_fl, grid_h=1)
            if rb['o'] is None:
                rb['o'] = _Readback(out, copy_stream)
            rb['o'].push(out)
            return imgs[0]
        if graphed is not None:
            img = graphed(**host_in)             # pinned host -> the graph's static input buffers (H2D on the compute stream), replay
        else:
            if stage['next'] is None:
                stage['next'] = upload()
            ts, ev = stage['next']
            cur.wait_event(ev)
            for t in ts.values():
                t.record_stream(cur)
            stage['next'] = upload()
            img = identity(ts)
        out = img if args.e2e_f32 else rt.layout_grid_u8(img, grid_w=T_C3, grid_h=1)
        if rb['o'] is None:
            rb['o'] = _Readback(out, copy_stream)
        rb['o'].push(out)
        return img

    with torch.no_grad():
        clocks = ClockSampler(local)      # started before the warm-up, see run_c2
        if rank == 0:
            clocks.start()
            time.sleep(1.0)
        for _ in range(max(args.warmup, 3)):
            step_resident()
        clocks.rows.clear()
        rt.reset_launch_count()
        ms = h.timed(step_resident, args.steps)
        launches = rt.launch_count() if graphed is None else launches_per_step * args.steps     # (a replay re-launches the captured kernels)
        for _ in range(2):
            step_e2e()
        ms_e2e = h.timed(step_e2e, args.steps)
        clk = clocks.stop() if rank == 0 else None

        roofline = breakdown = stages = None
        if not args.no_roofline:
            h.barrier()
            from invertavatar_b200 import triplane as _tp
            _tp.set_backbone_streams(False)
            side = rt.side_streams
            rt.side_streams = lambda device, n=2: (torch.cuda.current_stream(device),) * n    # one kernel at a time on the device
            try:
                rt.flop_count_begin()
                rt.profile_begin()
                for _ in range(args.steps):
                    identity(res_in)             # eager: per-launch events cannot be recorded inside a graph replay
                rep = rt.profile_report()
                counted = rt.flop_count_end()
            finally:
                rt.side_streams = side
                _tp.set_backbone_streams(None)
            h.barrier()
            if rank == 0:
                conv = rep.get('ia_conv_tc', {'ms': 0.0, 'launches': 0})
                roofline = _roofline(conv['ms'], conv['launches'], counted['algorithmic'], args.steps, {
                    'traffic': None,
                    'algorithmic_conv_flops_per_step': counted['algorithmic'] / args.steps,
                    'computed_conv_flops_per_step': counted['computed'] / args.steps,
                    'issued_mma_flops_per_step': counted['issued_mma'] / args.steps,
                    'issued_tflops': counted['issued_mma'] / (conv['ms'] * 1e-3) / 1e12 if conv['ms'] > 0 else 0.0,
                    'note': 'algorithmic = 2*MAC of every convolution of the step at true channel counts, strided encoder convolutions at '
                            'their output resolution, transposed convolutions at their input resolution (SURVEY 8d); computed = what the '
                            'device evaluates (stride-2 convolutions run at full resolution); issued = split terms x padded channels'})
                roofline['issued_frac'] = roofline['issued_tflops'] / roofline['peak']
                tot = sum(v['ms'] for v in rep.values())
                breakdown = {k: {'ms_per_step': v['ms'] / args.steps, 'launches_per_step': v['launches'] / args.steps, 'share': v['ms'] / tot}
                             for k, v in sorted(rep.items(), key=lambda kv: -kv[1]['ms'])}
            # wall-clock of the three stages of the step (device-timed, resident inputs)
            x = {'image': res_in['image'], 'uv': res_in['uv']}
            v = {'uvcoords_image': res_in['uvimg']}

            def t_ms(fn, n=3):
                fn(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    o = fn()
                e1.record(); torch.cuda.synchronize()
                return e0.elapsed_time(e1) / n, o
            enc_ms, ws = t_ms(lambda: net.encode(x['image'][:1]))
            tex = G.texture_backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
            sta = G.backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
            e4e = {'w': ws, 'texture': tex, 'static': sta}
            ar_ms, (upd, _) = t_ms(lambda: net.AR_eval_forward(x, res_in['c'], v, ws, [None, None], e4e_results=e4e, return_fake=False))
            fr_ms, _ = t_ms(lambda: G.synthesis_withTexture(ws, upd['texture'], res_in['c'][:1], {'uvcoords_image': res_in['uvimg'][:1]}, noise_mode='const',
                                                            static_feats=upd['static'], evaluation=True)['image'], n=8)
            frT_ms, _ = t_ms(lambda: G.synthesis_withTexture(ws.expand(T_C3, -1, -1), [f.expand(T_C3, -1, -1, -1) for f in upd['texture']], res_in['c'], v,
                                                             noise_mode='const', static_feats=[f.expand(T_C3, -1, -1, -1) for f in upd['static']],
                                                             evaluation=True)['image'], n=8)
            stages = {'encode_ms': enc_ms, 'ar_eval_forward_ms': ar_ms, 'synthesis_withTexture_ms_per_frame': fr_ms,
                      f'synthesis_withTexture_batch{T_C3}_ms': frT_ms, 'note': 'eager (un-graphed) stage times'}

        cpu = parity = None
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            fps, cores, times, ref_img = cpu_c3(args, 1, warmup=0)
            cpu = {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': f'1 identity step ({T_C3} frames) of the same workload, no warm-up (oracle port of the reference, torch CPU fp32, {cores} threads)'}
            jit_ar, u_ar, jit_frames = c3_draws(args)
            got = identity(res_in, draws=(jit_ar.to(dev), u_ar.to(dev), [j.to(dev) for j in jit_frames])).float().cpu()
            parity = _parity(got, ref_img, f'the {T_C3} driven frames of one identity step vs the oracle step of the cpu_baseline leg (same images, UV, cameras, pinned draws)')

    frames = T_C3 * n_fl * world * args.steps
    if rank != 0:
        return None
    if launches is not None:
        launches = launches * n_fl if graphed is not None else launches
    line = {'metric': METRIC_C3, 'value': frames / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args, world), 'clocks': clk,
            'e2e': {'value': frames / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': n_fl * sum(t.numel() * t.element_size() for t in host_in.values()),
                    'd2h_bytes_per_step': rb['o'].bytes, 'ms_per_step': ms_e2e / args.steps,
                    'readback': 'fp32 NCHW images' if args.e2e_f32 else 'uint8 HWC frames (layout_grid, eval_seq.py:214)'},
            'gpu_launches': launches, 'identities_per_s': n_fl * world * args.steps / (ms * 1e-3),
            'issue': 'eager (Python + ctypes launches)' if graphed is None else
                     'one CUDA graph per identity step (invertavatar_b200.graphs.GraphedCall), replayed' + (f'; {n_fl} identities in flight per step on {n_fl} streams' if n_fl > 1 else '')}
    if stages is not None:
        line['stages'] = stages
    if roofline is not None:
        line['roofline'] = roofline
        line['kernel_breakdown'] = breakdown
    if cpu is not None:
        line['cpu_baseline'] = cpu
        line['parity'] = parity
    return line


_JSON_OUT = None


def _emit_line(line):
    """The ONE JSON line goes to the process's original stdout; everything else written to fd 1 meanwhile (the NCCL version
    banner, library chatter) has been routed to stderr by main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    global _JSON_OUT
    args = parse_args()
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
