#!/usr/bin/env python
"""Headline benchmark: audio-seconds processed per wall-second (real-time factor) of the SOME inference
hot path (waveform -> log-mel -> two-head conformer -> decoded notes), BASELINE.json configs[1]:
configs/two_head_model.yaml, 64 x 30 s synthetic 44.1 kHz mono clips per GPU, bf16 operands / fp32 accumulate.

    python bench.py --gpus N --steps K --warmup W            # this repo (libsome_b200.so kernels)
    python bench.py --impl reference ...                     # the reference algorithm on the host CPU cores

A "step" = one pass of the hot path over one batch (64 clips x 30 s = 1920 audio-seconds per GPU; weak
scaling: each rank owns its own 64 clips, the only exchange is one all-gather of the packed notes).
  value  : device-timed (CUDA events, max over ranks), audio already resident in HBM when the region starts
  e2e    : the same metric through the plugin's public infer() with HOST numpy buffers: pinned H2D of the
           audio and D2H of the notes (+ the NCCL all-gather at N > 1) inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches: see the task contract.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from some_b200 import synth  # noqa: E402

CLIPS_PER_GPU = 64
CLIP_SECONDS = 30.0
CONFIG_NAME = 'two_head'
WORKLOAD = 'configs/two_head_model.yaml, batch=64x30 s synthetic 44.1 kHz mono clips per GPU, bf16 operands'
METRIC = 'audio-seconds/sec (real-time factor) two_head conformer'


def load_peaks():
    try:
        with open(os.path.join(REPO, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'tf_burst': p['bf16_tflops'], 'tf_sustained': p['bf16_tflops_sustained'],
                'source': 'measured (MEASURED_PEAKS.json)'}
    except Exception:
        return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


def make_clips(first_index, count, seconds):
    """Deterministic synthetic clips.  Generating 64 x 30 s sung-note signals costs ~1 s each; to keep the
    default run within minutes a pool of 8 distinct clips is generated and rotated with distinct gains and
    circular shifts (still 64 different waveforms of the named length)."""
    n = int(round(seconds * synth.SR))
    pool = [synth.synth_waveform(1000 + j, num_samples=n) for j in range(min(8, count))]
    clips = []
    for i in range(count):
        gi = first_index + i
        base = pool[gi % len(pool)]
        clips.append(np.ascontiguousarray(np.roll(base, 4099 * (gi // len(pool))) * np.float32(1.0 - 0.01 * (gi % 7))))
    return clips


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '50', '-i', str(self.gpu_index)],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                out['sm_max_mhz'] = float(f[2])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out['sm_mhz'] = float(np.median(sm))
            out['samples'] = len(sm)
        out['reasons'] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------- reference arm
def time_cpu_reference(clips, threads=None):
    """Times the reference ALGORITHM on the host cores: the oracle port (oracle/decode.infer = the serial
    batch-1 loop of inference/base_infer.py:46-53 in fp32 torch).  /root/reference itself is not on the GPU box;
    the port is pinned to it by tests/golden."""
    from oracle import decode as odecode
    if threads:
        torch.set_num_threads(threads)
    config = synth.named_config(CONFIG_NAME)
    sd = synth.fabricate_state_dict(config, seed=1234)
    t0 = time.perf_counter()
    odecode.infer(sd, config, clips, quantized=CONFIG_NAME.startswith('quant'))
    dt = time.perf_counter() - t0
    return sum(len(c) for c in clips) / synth.SR / dt, dt


def pick_cpu_threads(clip):
    """The reference's small fp32 GEMMs do not scale to every core of a big host (64 threads were slower than 8 in
    the first measurements), and torchrun pins OMP_NUM_THREADS=1.  Be fair to the CPU arm: try a few thread counts
    on one clip and keep the fastest; the count used is reported as `cores`."""
    avail = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    cands = sorted({c for c in (8, 16, 32, avail) if c <= avail} | {avail})
    best, best_v = avail, -1.0
    for c in cands:
        time_cpu_reference([clip], threads=c)                # warm-up at this thread count
        v, _ = time_cpu_reference([clip], threads=c)
        if v > best_v:
            best, best_v = c, v
    torch.set_num_threads(best)
    return best


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_clips = args.ref_clips
    clips = make_clips(0, n_clips, CLIP_SECONDS)
    cores = pick_cpu_threads(clips[0])
    for _ in range(args.warmup):
        time_cpu_reference(clips[:1])
    times = []
    for _ in range(args.steps):
        _, dt = time_cpu_reference(clips)
        times.append(dt)
    total = n_clips * CLIP_SECONDS * args.steps / sum(times)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': total, 'unit': 'audio-s/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * sum(times) / len(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': f'{n_clips} x {CLIP_SECONDS:.0f} s clips per step (serial batch-1 loop)'},
        'cpu_baseline': {'value': total, 'unit': 'audio-s/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{n_clips} x {CLIP_SECONDS:.0f} s clips x {args.steps} steps, torch fp32, {cores} threads'},
        'e2e': {'value': total, 'unit': 'audio-s/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- this repo
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from some_b200 import dist as sdist
    from some_b200 import plugin

    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    config = synth.named_config(CONFIG_NAME)
    import contextlib
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(sys.stderr):   # stdout = the ONE JSON line
        ckpt = synth.write_checkpoint(d, config, seed=1234)
        cls = plugin.QuantizedMIDIExtractionInference if CONFIG_NAME.startswith('quant') else plugin.MIDIExtractionInference
        ins = cls(config=config, model_path=ckpt, device=f'cuda:{local_rank}')
    eng = ins.model
    quant = CONFIG_NAME.startswith('quant')
    clips = make_clips(rank * CLIPS_PER_GPU, CLIPS_PER_GPU, CLIP_SECONDS)
    audio_seconds_rank = sum(len(c) for c in clips) / synth.SR
    lengths_all = [len(c) for c in clips] * world           # every rank's clips have the same lengths
    shards = [list(range(r * CLIPS_PER_GPU, (r + 1) * CLIPS_PER_GPU)) for r in range(world)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident arm ("value")
    host, tables, cu = eng.pack(clips)
    b, m, max_frames = len(clips), int(cu[-1]), int(np.diff(cu).max())
    wave = host.to(dev)
    tables_d, cu_d = tables.to(dev), torch.from_numpy(cu).to(dev)
    ws = eng.workspace(m)
    note_count = torch.empty(b, dtype=torch.int32, device=dev)

    def device_step():
        eng.run_mel(wave, tables_d[:b], tables_d[b:], cu_d, b, max_frames, None, ws.units)
        eng.run_trunk(ws, m, b, cu_d, max_frames, 'softmax' if quant else 'sigmoid')
        eng.run_decode(ws, m, b, cu_d, note_count, quant)

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        device_step()                       # product path: native launch sequencer (some_forward)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launches - launches0
    # second pass of the same K steps with CUDA events around every launch (per-kernel Python path) for the roofline
    eng.start_profile(cu)
    barrier()
    for _ in range(args.steps):
        device_step()
    prof = eng.stop_profile()
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- end-to-end arm ("e2e"): host numpy in, host numpy out, through the plugin
    # The step's inputs sit in PINNED host memory (numpy views of page-locked buffers, as a production loader would hand
    # them over): the engine copies host -> device straight from them.  (Pageable numpy arrays go through a pinned staging
    # memcpy first; that variant is reported as e2e.pageable_value.)
    from some_b200.engine import pinned_array
    pinned_clips = []
    for c in clips:
        a = pinned_array(len(c))
        a[:] = c
        pinned_clips.append(a)
    all_clips = None
    if world > 1:
        # every rank needs the LENGTHS of all clips (they are identical here); only its own shard's samples are touched
        all_clips = [pinned_clips[i % CLIPS_PER_GPU] for i in range(world * CLIPS_PER_GPU)]

    def e2e_step(src=None):
        if world > 1:
            return sdist.infer_sharded(ins, all_clips)      # shard -> infer -> ONE NCCL all-gather of the packed notes
        return ins.infer(pinned_clips if src is None else src)

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    pageable_s = None
    if world == 1:                                  # same call with ordinary (pageable) numpy inputs
        e2e_step(clips)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(min(args.steps, 3)):
            e2e_step(clips)
        pageable_s = (time.perf_counter() - t1) / min(args.steps, 3)
    h2d = int(host.numel() * 4 + tables.numel() * 8 + cu.nbytes)
    d2h = int(m * 9 + b * 4)

    # ---------------- max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1000.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return
    total_audio = audio_seconds_rank * world * args.steps
    value = total_audio / (dev_ms / 1000.0)
    e2e_value = total_audio / (e2e_ms / 1000.0)

    peaks = load_peaks()
    gemm = prof.get('some_gemm', {'ms': 0.0, 'work': 0.0, 'launches': 0})
    # dominant kernel = K-gemm (all shapes of a step aggregated): achieved = algorithmic FLOPs / event time
    achieved_tf = gemm['work'] / (gemm['ms'] / 1000.0) / 1e12 if gemm['ms'] > 0 else 0.0
    peak_tf = peaks['tf_sustained']                        # kernel timed inside a long step
    kernels = {}
    step_kernel_ms = sum(v['ms'] for v in prof.values())
    for name, v in prof.items():
        rate = v['work'] / (v['ms'] / 1000.0) if v['ms'] > 0 else 0.0
        tensor = name in ('some_gemm', 'some_attention_varlen')
        kernels[name] = {'launches_per_step': v['launches'] // args.steps, 'ms_per_step': v['ms'] / args.steps,
                         'share': v['ms'] / step_kernel_ms if step_kernel_ms else 0.0,
                         ('tflops' if tensor else 'gbs'): rate / (1e12 if tensor else 1e9),
                         'frac': rate / ((peak_tf * 1e12) if tensor else (peaks['hbm_gbs'] * 1e9))}
    traffic = None
    try:
        with open(os.path.join(REPO, 'profiles', 'r01_gemm_traffic.json')) as f:
            traffic = json.load(f).get('dram_bytes_per_launch')
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ref_clips = make_clips(0, args.ref_clips, CLIP_SECONDS)
        cores = pick_cpu_threads(ref_clips[0])
        v, dt = time_cpu_reference(ref_clips)
        cpu = {'value': v, 'unit': 'audio-s/s', 'cores': cores, 'kind': 'port',
               'sample': f'{args.ref_clips} x {CLIP_SECONDS:.0f} s clips of the same workload, oracle port '
                         f'(torch fp32, serial batch-1 loop), {dt:.1f} s of CPU work'}

    line = {
        'metric': METRIC, 'value': value, 'unit': 'audio-s/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'clips_per_gpu': CLIPS_PER_GPU, 'clip_seconds': CLIP_SECONDS,
                   'frames_per_gpu': m, 'parallelism': f'dp{world}', 'l2': f'inputs ({host.numel() * 4 / 1e6:.0f} MB audio, {m * 512 * 4 * 2 / 1e9:.1f} GB residual streams) exceed the 126 MB L2',
                   'weights': 'seeded random (no pretrained checkpoint offline)'},
        'e2e': {'value': e2e_value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_ms / args.steps, 'inputs': 'pinned host numpy arrays -> plugin.infer -> host numpy notes',
                'pageable_value': (audio_seconds_rank / pageable_s) if pageable_s else None},
        'gpu_launches': launches,
        'roofline': {'kernel': 'some_gemm (tcgen05, all shapes of a step)', 'bound': 'tensor', 'achieved': achieved_tf,
                     'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved_tf / peak_tf if peak_tf else None,
                     'traffic': traffic, 'peak_source': peaks['source'] + ', sustained bf16',
                     'measured_in': 'second pass of the same K steps with CUDA events around every launch'},
        'kernels': kernels,
        'clocks': clocks,
    }
    if cpu is not None:
        line['cpu_baseline'] = cpu
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--ref-clips', type=int, default=8, help='bounded CPU sample: clips per CPU step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    # extra measurements for the BASELINE.md table (the default = the contract workload, BASELINE.json configs[1])
    ap.add_argument('--config', default='two_head', choices=['two_head', 'quant_two_head', 'midi_conformer'])
    ap.add_argument('--clips', type=int, default=64, help='clips per GPU')
    ap.add_argument('--seconds', type=float, default=30.0, help='clip length')
    args = ap.parse_args()
    global CONFIG_NAME, CLIPS_PER_GPU, CLIP_SECONDS, WORKLOAD
    CONFIG_NAME, CLIPS_PER_GPU, CLIP_SECONDS = args.config, args.clips, args.seconds
    yaml_name = {'two_head': 'two_head_model', 'quant_two_head': 'quant_two_head_model', 'midi_conformer': 'midi_conformer'}[args.config]
    WORKLOAD = (f'configs/{yaml_name}.yaml, batch={args.clips}x{args.seconds:g} s synthetic 44.1 kHz mono clips per GPU, '
                f'bf16 operands')
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
