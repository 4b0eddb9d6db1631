#!/usr/bin/env python
"""Benchmark of the InvertAvatar generator-forward hot path (BASELINE.json metric: 512^2 avatar frames/s at
128^2 neural render x 48+48 depth samples).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

A "step" renders one batch of `--batch` (default 8 = BASELINE configs[1]) frames per GPU: mapping (z,c -> ws) +
TriPlaneGenerator.synthesis (3 StyleGAN2 backbones, UV rasterize/stitch, fused volume renderer, super-resolution) on
synthetic latents / cameras / UV mesh conditions and random-init weights of the reference architecture.  Frames are
independent, so N GPUs run N x batch frames (weak scaling); the only collective is the all-gather of the final images.

One JSON line on stdout (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same through the public
API with HOST inputs (pinned H2D of z, c, uv every step) and a D2H copy of the images every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = '512x512 avatar frames/sec (128x128 neural render x 48+48 depth samples)'
UNIT = 'frames/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=8, help='frames per GPU per step')
    ap.add_argument('--res', type=int, default=128, help='neural rendering resolution')
    ap.add_argument('--depth', type=int, default=48, help='coarse = importance depth samples per ray')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {'workload': f'Next3D++ reenactment 512^2, {args.res}^2 neural x {args.depth}+{args.depth} depth, batch {args.batch}/GPU '
                        f'(BASELINE configs[1]), random-init generator, synthetic latents/cameras/UV',
            'frames_per_gpu_per_step': args.batch, 'global_frames_per_step': args.batch * n_gpus,
            'neural_res': args.res, 'depth_samples': [args.depth, args.depth], 'parallelism': f'dp{n_gpus} (frames sharded, weights replicated)',
            'conv_precision': 'bf16x3 split operands (hi*hi+hi*lo+lo*hi), fp32 accumulate in TMEM', 'gather': getattr(args, 'gather_kind', 'none'),
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
        elif isinstance(m, sg.ToRGBLayer):
            pass
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
        return {'sm_mhz': (sm[len(sm) // 2] if sm else None), 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated in oracle/ (the reference is Python/torch and is not present on the GPU
# box; the oracle is pinned to it by tests/golden).  Used as cpu_baseline of the B200 arm and as --impl reference.
# ----------------------------------------------------------------------------------------------------------------
def cpu_frames_per_s(args, frames, repeats, warmup=1):
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
    with torch.no_grad():
        for _ in range(warmup):
            step()
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            step()
            times.append(time.perf_counter() - t0)
    return frames * len(times) / sum(times), cores, times


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    frames = 1   # bounded sample: one frame of the same workload per step (the reference's throughput is flat in batch)
    steps = max(1, min(args.steps, 5))
    fps, cores, times = cpu_frames_per_s(args, frames, steps, warmup=max(1, min(args.warmup, 1)))
    sample = f'{frames} frame/step x {steps} steps of the same workload (reference algorithm = oracle port, torch CPU fp32, {cores} threads)'
    line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps, 'warmup': 1,
            'ms_per_step': 1000.0 * sum(times) / len(times), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args, args.gpus),
            'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    _emit_line(line)


# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from invertavatar_b200 import _C, synth
    from invertavatar_b200 import runtime as rt
    from invertavatar_b200.triplane import TriPlaneGenerator

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
    img_h = torch.empty((B, 3, 512, 512), dtype=torch.float32).pin_memory()
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
                                                                      'stores to peer-mapped symmetric memory') + ' + barrier'
            except Exception as ex:   # symmetric memory unavailable on this box / build
                peer = None
                gather_kind += f' (symmetric memory unavailable: {type(ex).__name__})'

    def frame_batch(z, cond, c, uv):
        ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
        if peer is not None:
            with peer.sink():
                img = G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=res, noise_mode='const', evaluation=True)['image']
            peer.barrier()
            return img
        img = G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=res, noise_mode='const', evaluation=True)['image']
        if world > 1:
            dist.all_gather_into_tensor(gathered, img.contiguous())
        return img

    def step_resident():
        return frame_batch(z, cond, c, uv)

    # End-to-end step: pinned-host inputs are uploaded on the compute stream every step; the images go back to pinned host
    # memory on a second stream (double-buffered), so the device->host read of step i overlaps the kernels of step i+1.
    # Every copy is enqueued inside the timed region and the closing synchronize waits for all of them.
    copy_stream = torch.cuda.Stream(device=dev)
    img_hs = [img_h, torch.empty_like(img_h).pin_memory()]
    e2e_state = {'i': 0, 'done': [None, None]}

    def step_e2e():
        zz, cc, c2, uu = z_h.to(dev, non_blocking=True), cond_h.to(dev, non_blocking=True), c_h.to(dev, non_blocking=True), uv_h.to(dev, non_blocking=True)
        img = frame_batch(zz, cc, c2, uu)
        k = e2e_state['i'] & 1
        e2e_state['i'] += 1
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            if e2e_state['done'][k] is not None:
                e2e_state['done'][k].synchronize()      # the host buffer's previous read-back has landed (2 steps ago)
            img_hs[k].copy_(img, non_blocking=True)
            img.record_stream(copy_stream)
            done = torch.cuda.Event()
            done.record(copy_stream)
            e2e_state['done'][k] = done
        return img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        torch.cuda.current_stream().wait_stream(copy_stream)   # the last read-back belongs to the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step_resident()
        if peer is not None:   # the fused gather must reproduce the NCCL gather bit for bit
            img = step_resident()
            dist.all_gather_into_tensor(gathered, img.contiguous())
            torch.cuda.synchronize()
            ok = torch.tensor([1 if torch.equal(gathered, peer.tensor) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            assert int(ok.item()) == 1, 'fused peer-memory gather differs from the NCCL all-gather'
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()   # samples every 200 ms across both timed regions (resident + end-to-end)
        rt.reset_launch_count()
        ms = timed(step_resident, args.steps)
        launches = rt.launch_count()
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        clk = clocks.stop() if rank == 0 else None

        roofline = None
        breakdown = None
        if not args.no_roofline:
            # same steps again with every launch bracketed by CUDA events on the launching stream (every rank runs them so
            # that the image all-gather stays matched; rank 0 reports)
            barrier()
            from invertavatar_b200 import triplane as _tp
            _tp.set_backbone_streams(False)      # per-kernel durations are taken with one kernel at a time on the device
            rt.profile_begin()
            for _ in range(args.steps):
                step_resident()
            rep = rt.profile_report()
            _tp.set_backbone_streams(None)
            barrier()
        if rank == 0 and not args.no_roofline:
            conv = rep.get('ia_conv_tc', {'ms': 0.0, 'launches': 0})
            flops = conv_flops_per_frame(G) * B * args.steps
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
            except Exception:
                pass
            peak = float(peaks.get('bf16_tflops_sustained', 1400.0))
            achieved = flops / (conv['ms'] * 1e-3) / 1e12 if conv['ms'] > 0 else 0.0
            roofline = {'kernel': 'conv_tc_kernel (tcgen05 implicit-GEMM modulated convolution)', 'bound': 'tensor', 'achieved': achieved,
                        'peak': peak, 'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peaks else 'fallback 1.4 PFLOP/s sustained (of fallback)',
                        'unit': 'TFLOP/s', 'frac': achieved / peak,
                        # DRAM bytes of one launch from the committed `ncu --set full` capture (profiles/r1_conv_same_full_v4.txt):
                        # the 128->128 @512^2 layer at batch 8 reads 1.085 GB and writes 1.030 GB; its algorithmic bytes are the
                        # bf16 hi/lo operand in (1.074 GB) and the next layer's operand out (1.074 GB) -- no re-reads
                        'traffic': 2.115e9, 'traffic_launch': 'conv_tc2 128->128 @512x512, batch 8 (algorithmic 2.147e9 B)',
                        'algorithmic_flops_per_frame': conv_flops_per_frame(G), 'launches_per_step': conv['launches'] / args.steps,
                        'avg_launch_ms': conv['ms'] / max(1, conv['launches']),
                        'issued_tflops': 3.0 * achieved, 'issued_frac': 3.0 * achieved / peak,
                        'note': 'achieved/frac count algorithmic FLOPs (fp32 semantics); the 3-term bf16 split that the 1e-3 parity bar '
                                'requires issues 3x these MMAs (issued_*), so frac <= 1/3 by construction'}
            tot = sum(v['ms'] for v in rep.values())
            breakdown = {k: {'ms_per_step': v['ms'] / args.steps, 'launches_per_step': v['launches'] / args.steps, 'share': v['ms'] / tot}
                         for k, v in sorted(rep.items(), key=lambda kv: -kv[1]['ms'])}

    frames = B * world * args.steps
    value = frames / (ms * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in (z_h, cond_h, c_h, uv_h))
    d2h = img_h.numel() * img_h.element_size()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    args.gather_kind = gather_kind
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args, world), 'clocks': clk,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches}
    if roofline is not None:
        line['roofline'] = roofline
        line['kernel_breakdown'] = breakdown
    if world == 1 and not args.no_cpu_baseline:
        fps, cores, times = cpu_frames_per_s(args, 1, 2, warmup=1)
        line['cpu_baseline'] = {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': f'1 frame/step x 2 steps of the same workload after 1 warm-up (oracle port of the reference, torch CPU fp32, {cores} threads)'}
    _emit_line(line)


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
