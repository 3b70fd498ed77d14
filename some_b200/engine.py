"""Batched SOME inference engine: packs clips var-len, launches the sm_100a kernels of
libsome_b200.so in order on the current CUDA stream and unpacks the decoded notes.

Equivalent to running the reference's batch-1 loop (inference/base_infer.py:46-53) once per clip:
clips never interact (per-clip attention, per-clip zero-padded depthwise conv, per-clip decode).
Launch sequence per conform_blocke (Gconform.py:56-63), both streams (midi / bound) in every launch:
    LN1 -> GEMM(ffn1.ln1)+SiLU -> GEMM(ffn1.ln2)*0.5+x -> LN2 -> GEMM(to_q|to_kv) -> attention ->
    GEMM(to_out)+x -> LN3 -> GEMM(pointwise_conv1)+GLU -> dwconv+BN+SiLU -> GEMM(pointwise_conv2)+x ->
    LN4 -> GEMM(ffn2.ln1)+SiLU -> GEMM(ffn2.ln2)*0.5+x -> LN5
The residual stream x is fp32 [M, 512]; GEMM operands are bf16; accumulation is fp32.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .config import DIM, FFN_DIM, check_supported
from .weights import ModelWeights, build_c_model, mel_tables

HOP = 512


def frames_of(num_samples: int) -> int:
    return 1 + num_samples // HOP


def pinned_array(num_samples: int) -> np.ndarray:
    """float32 numpy array in page-locked host memory.  A loader that decodes audio into such buffers lets Engine.infer
    copy host -> device straight from them (no staging memcpy); ordinary numpy arrays work too, through a staging copy."""
    return torch.empty(int(num_samples), dtype=torch.float32).pin_memory().numpy()


class _Workspace:
    def __init__(self, m: int, outdim: int, device):
        bf, f32 = torch.bfloat16, torch.float32
        self.m = m
        self.x = torch.empty((2, m, DIM), dtype=f32, device=device)          # residual streams
        self.a = torch.empty((2, m, DIM), dtype=bf, device=device)           # LN out / attention out / dwconv out
        self.h = torch.empty((2, m, FFN_DIM), dtype=bf, device=device)       # FFN hidden
        self.qkv = torch.empty((2, m, 3 * DIM), dtype=bf, device=device)
        self.g = torch.empty((2, m, DIM), dtype=bf, device=device)           # GLU out (dwconv in)
        self.units = torch.empty((m, 80), dtype=bf, device=device)
        self.probs = torch.empty((m, outdim), dtype=f32, device=device)
        self.bounds = torch.empty((m,), dtype=f32, device=device)
        self.note_midi = torch.empty((m,), dtype=f32, device=device)
        self.note_dur = torch.empty((m,), dtype=torch.int32, device=device)
        self.note_rest = torch.empty((m,), dtype=torch.uint8, device=device)
        self.scratch = torch.empty((int(_lib.load().some_decode_scratch_bytes(m)),), dtype=torch.uint8, device=device)
        c = self.c = _lib.WorkspaceC()                    # some_workspace for the native sequencer
        for s in range(2):
            c.x[s], c.a[s], c.h[s] = self.x[s].data_ptr(), self.a[s].data_ptr(), self.h[s].data_ptr()
            c.qkv[s], c.g[s] = self.qkv[s].data_ptr(), self.g[s].data_ptr()
        c.units, c.probs, c.bounds = self.units.data_ptr(), self.probs.data_ptr(), self.bounds.data_ptr()


class Engine:
    def __init__(self, config: dict, state_dict, device='cuda'):
        self.lib = _lib.load()
        check_supported(config)
        self.config = config
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.SomeB200Error('some_b200 runs on CUDA devices only (sm_100a); there is no CPU path')
        self.quantized = False
        self.w = ModelWeights(state_dict, config, self.device)
        self._cmodel, self._cmodel_keep = build_c_model(self.w)
        self.mel = mel_tables(config, self.device)
        self.outdim = config['midi_num_bins']
        self.timestep = config['hop_size'] / config['audio_sample_rate']
        self._ws: Optional[_Workspace] = None
        self.launches = 0
        self._sum_t2 = 0.0
        # launches of one trunk pass: inln + (lay + 1) blocks x 15 + lay GLU mixes + (final LN + bound head - LN5) + head
        self.trunk_launches = 1 + 15 * (self.w.lay + 1) + self.w.lay + 1 + 1
        # optional per-kernel timing: name -> [(start_event, end_event, work)] where work = FLOPs (GEMM,
        # attention) or algorithmic bytes (HBM-bound kernels); enabled by bench.py via start_profile()
        self.prof: Optional[dict] = None

    def start_profile(self, cu_frames_host=None):
        self.prof = {}
        if cu_frames_host is not None:
            t = np.diff(np.asarray(cu_frames_host)).astype(np.float64)
            self._sum_t2 = float((t * t).sum())

    def stop_profile(self) -> Dict[str, dict]:
        """Returns {kernel: {launches, ms, work}} from the CUDA events recorded since start_profile()."""
        torch.cuda.synchronize(self.device)
        out = {}
        for name, recs in (self.prof or {}).items():
            out[name] = {'launches': len(recs), 'ms': float(sum(a.elapsed_time(b) for a, b, _ in recs)),
                         'work': float(sum(w for _, _, w in recs))}
        self.prof = None
        return out

    def _mark(self, name: str, work: float):
        """Context manager: CUDA events on the launching stream around one kernel launch."""
        eng = self

        class _M:
            def __enter__(self_m):
                if eng.prof is not None:
                    self_m.a = torch.cuda.Event(enable_timing=True)
                    self_m.b = torch.cuda.Event(enable_timing=True)
                    self_m.a.record(torch.cuda.current_stream(eng.device))
                return self_m

            def __exit__(self_m, *exc):
                if eng.prof is not None:
                    self_m.b.record(torch.cuda.current_stream(eng.device))
                    eng.prof.setdefault(name, []).append((self_m.a, self_m.b, work))
                eng.launches += 1
                return False

        return _M()

    # ------------------------------------------------------------------ helpers
    @property
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def workspace(self, m: int) -> _Workspace:
        if self._ws is None or self._ws.m < m:
            self._ws = None
            self._ws = _Workspace(max(m, 1), self.outdim, self.device)
        return self._ws

    def _gemm(self, a0, a1, w0, w1, b0, b1, out0, out1, r0, r1, m, n, k, lda, ld_out, epi, alpha=1.0, groups=2):
        g = _lib.GemmArgs()
        g.A, g.W = _lib.pair(a0, a1), _lib.pair(w0, w1)
        g.bias, g.out, g.resid = _lib.pair(b0, b1), _lib.pair(out0, out1), _lib.pair(r0, r1)
        g.groups, g.M, g.N, g.K, g.lda, g.ld_out, g.epilogue, g.alpha = groups, m, n, k, lda, ld_out, epi, alpha
        with self._mark('some_gemm', 2.0 * m * n * k * groups):
            _lib.check(self.lib.some_gemm(C.byref(g), self._stream), 'some_gemm')

    def _ln(self, x, gamma, beta, out_bf16, out_f32, m):
        a = _lib.LnArgs()
        a.x = _lib.pair(x[0], x[1])
        a.gamma, a.beta = _lib.pair(*gamma), _lib.pair(*beta)
        a.out_bf16 = _lib.pair(out_bf16[0], out_bf16[1]) if out_bf16 is not None else (C.c_void_p * 2)()
        a.out_f32 = _lib.pair(out_f32[0], out_f32[1]) if out_f32 is not None else (C.c_void_p * 2)()
        a.groups, a.M = 2, m
        nout = (2 if out_bf16 is not None else 0) + (4 if out_f32 is not None else 0)
        with self._mark('some_layernorm', 2.0 * m * DIM * (4 + nout)):
            _lib.check(self.lib.some_layernorm(C.byref(a), self._stream), 'some_layernorm')

    # ------------------------------------------------------------------ stages
    def run_mel(self, wave: torch.Tensor, clip_start: torch.Tensor, clip_len: torch.Tensor, cu_frames: torch.Tensor,
                b: int, max_frames: int, out_f32: Optional[torch.Tensor], out_bf16: Optional[torch.Tensor]):
        """K-mel (spec.py:38-72).  wave f32; clip i = wave[start_i : start_i + len_i]; outputs [M, 80]."""
        t = self.mel
        m = (out_f32 if out_f32 is not None else out_bf16).shape[0]
        with self._mark('some_mel_logmel', m * (512 * 4 + 80 * 4.0)):        # 2368 B / frame
            _lib.check(self.lib.some_mel_logmel(
                wave.data_ptr(), clip_start.data_ptr(), clip_len.data_ptr(), cu_frames.data_ptr(), b, max_frames,
                t['mel_start'].data_ptr(), t['mel_count'].data_ptr(), t['mel_weights'].data_ptr(),
                t['twiddle'].data_ptr(), t['window'].data_ptr(), _lib.ptr(out_f32), _lib.ptr(out_bf16),
                1e-5, self._stream), 'some_mel_logmel')

    def _block(self, ws: _Workspace, blk, m: int, b: int, cu_frames, max_frames: int, last: bool):
        x, a, h, qkv, g = ws.x, ws.a, ws.h, ws.qkv, ws.g
        lib, st = self.lib, self._stream
        w0, w1 = blk

        def ln(i, out_bf16=a, out_f32=None):
            self._ln(x, (w0.ln_g[i], w1.ln_g[i]), (w0.ln_b[i], w1.ln_b[i]), out_bf16, out_f32, m)

        def ffn(i):
            f0, f1 = w0.ffn[i], w1.ffn[i]
            self._gemm(a[0], a[1], f0['w1'], f1['w1'], f0['b1'], f1['b1'], h[0], h[1], None, None,
                       m, FFN_DIM, DIM, DIM, FFN_DIM, _lib.EPI_SILU_BF16)
            self._gemm(h[0], h[1], f0['w2'], f1['w2'], f0['b2'], f1['b2'], x[0], x[1], x[0], x[1],
                       m, DIM, FFN_DIM, FFN_DIM, DIM, _lib.EPI_RESID_F32, alpha=0.5)

        ln(0)
        ffn(0)                                                                   # Gconform.py:57
        ln(1)
        self._gemm(a[0], a[1], w0.w_qkv, w1.w_qkv, None, None, qkv[0], qkv[1], None, None,
                   m, 3 * DIM, DIM, DIM, 3 * DIM, _lib.EPI_STORE_BF16)
        at = _lib.AttnArgs()
        at.qkv, at.out = _lib.pair(qkv[0], qkv[1]), _lib.pair(a[0], a[1])
        at.groups, at.B, at.M, at.cu_frames, at.max_frames = 2, b, m, cu_frames.data_ptr(), max_frames
        with self._mark('some_attention_varlen', 2.0 * self._att_flops):
            _lib.check(lib.some_attention_varlen(C.byref(at), st), 'some_attention_varlen')
        self._gemm(a[0], a[1], w0.w_out, w1.w_out, w0.b_out, w1.b_out, x[0], x[1], x[0], x[1],
                   m, DIM, DIM, DIM, DIM, _lib.EPI_RESID_F32)                   # :60
        ln(2)
        self._gemm(a[0], a[1], w0.w_pw1, w1.w_pw1, w0.b_pw1, w1.b_pw1, g[0], g[1], None, None,
                   m, 2 * DIM, DIM, DIM, DIM, _lib.EPI_GLU_BF16)                # base_conv.py:65
        dw = _lib.DwconvArgs()
        dw.x, dw.w, dw.b = _lib.pair(g[0], g[1]), _lib.pair(w0.w_dw, w1.w_dw), _lib.pair(w0.b_dw, w1.b_dw)
        dw.out = _lib.pair(a[0], a[1])
        dw.groups, dw.B, dw.cu_frames, dw.max_frames = 2, b, cu_frames.data_ptr(), max_frames
        with self._mark('some_dwconv_bn_silu', 2.0 * m * DIM * 4):               # bf16 in + bf16 out
            _lib.check(lib.some_dwconv_bn_silu(C.byref(dw), st), 'some_dwconv_bn_silu')   # base_conv.py:66-68
        self._gemm(a[0], a[1], w0.w_pw2, w1.w_pw2, w0.b_pw2, w1.b_pw2, x[0], x[1], x[0], x[1],
                   m, DIM, DIM, DIM, DIM, _lib.EPI_RESID_F32)                   # base_conv.py:69 + Gconform.py:61
        ln(3)
        ffn(1)                                                                   # :62
        if not last:
            ln(4, out_bf16=a, out_f32=x)                                         # :63 (residual for the next Gcf)
        else:
            # final pair: midi stream -> normalised bf16 for outln; bound stream -> fused norm5 + cutheard + sigmoid
            al = _lib.LnArgs()
            al.x, al.gamma, al.beta = _lib.pair(x[0]), _lib.pair(w0.ln_g[4]), _lib.pair(w0.ln_b[4])
            al.out_bf16, al.out_f32 = _lib.pair(a[0]), (C.c_void_p * 2)()
            al.groups, al.M = 1, m
            with self._mark('some_layernorm', m * DIM * 6.0):
                _lib.check(lib.some_layernorm(C.byref(al), st), 'some_layernorm')
            with self._mark('some_bound_head', m * DIM * 4.0):
                _lib.check(lib.some_bound_head(x[1].data_ptr(), w1.ln_g[4].data_ptr(), w1.ln_b[4].data_ptr(),
                                               self.w.w_cut.data_ptr(), self.w.b_cut, m, ws.bounds.data_ptr(), st),
                           'some_bound_head')

    def run_trunk(self, ws: _Workspace, m: int, b: int, cu_frames: torch.Tensor, max_frames: int,
                  head: str = 'sigmoid', taps: Optional[dict] = None):
        """Gmidi_conform.forward (Gconform.py:119-140) + the head activation of midi_conforms.forward
        (Gmidi_conform.py:30-40).  Reads ws.units; writes ws.probs [m, outdim] and ws.bounds [m].
        head: 'sigmoid' | 'softmax' | 'logits'."""
        w, x, a = self.w, ws.x, ws.a
        epi = {'sigmoid': _lib.EPI_SIGMOID_F32, 'softmax': _lib.EPI_SOFTMAX_F32, 'logits': _lib.EPI_BIAS_F32}[head]
        if self.prof is None and taps is None:
            # product path: one native call enqueues the whole launch sequence (csrc/forward.cu)
            _lib.check(self.lib.some_forward(C.byref(self._cmodel), C.byref(ws.c), m, b, cu_frames.data_ptr(), max_frames,
                                             epi, self._stream), 'some_forward')
            self.launches += self.trunk_launches
            return
        # per-kernel path (CUDA events around every launch / intermediate taps): same sequence, driven from Python
        # QK^T + PV MACs of one attention launch (both streams): 2 * 8 heads * 64 * sum T^2 (profiling only; the clip
        # lengths come from the host copy of cu_frames so that no device sync sneaks into the timed region)
        self._att_flops = float(2 * 2 * 512 * self._sum_t2) if self.prof is not None else 0.0
        self._gemm(ws.units, ws.units, w.w_in[0], w.w_in[1], w.b_in[0], w.b_in[1], x[0], x[1], None, None,
                   m, DIM, 80, 80, DIM, _lib.EPI_BIAS_F32)                       # inln / inln1
        for i in range(w.lay):
            self._block(ws, w.blocks[i], m, b, cu_frames, max_frames, last=False)
            # Gcf.forward :85-87: midi += GLU(glu2(bound)); bound += GLU(glu1(midi))  (a = bf16 copies of norm5 out)
            self._gemm(a[1], a[0], w.glu_w[i][1], w.glu_w[i][0], w.glu_b[i][1], w.glu_b[i][0],
                       x[0], x[1], x[0], x[1], m, 2 * DIM, DIM, DIM, DIM, _lib.EPI_GLU_RESID_F32)
            if taps is not None:
                taps[f'model.cf_lay.{i}:midi'] = x[0, :m].clone()
                taps[f'model.cf_lay.{i}:bound'] = x[1, :m].clone()
        self._block(ws, w.blocks[w.lay], m, b, cu_frames, max_frames, last=True)
        self._gemm(a[0], None, w.w_head, None, w.b_head, None, ws.probs, None, None, None,
                   m, self.outdim, DIM, DIM, self.outdim, epi, groups=1)          # outln (+ sigmoid / softmax)

    def run_decode(self, ws: _Workspace, m: int, b: int, cu_frames: torch.Tensor, note_count: torch.Tensor,
                   quantized: bool, dbg: Optional[dict] = None, probs=None, bounds=None, out=None):
        cfg = self.config
        d = _lib.DecodeArgs()
        d.probs = (probs if probs is not None else ws.probs).data_ptr()
        d.bounds = (bounds if bounds is not None else ws.bounds).data_ptr()
        d.cu_frames = cu_frames.data_ptr()
        d.B, d.M, d.N, d.quantized = b, m, self.outdim, int(quantized)
        d.vmin, d.vmax = float(cfg['midi_min']), float(cfg['midi_max'])
        d.deviation = float(cfg.get('midi_prob_deviation', 1.0))
        d.threshold = float(cfg.get('rest_threshold', 0.1))
        nm_t, nd_t, nr_t = out if out is not None else (ws.note_midi, ws.note_dur, ws.note_rest)
        d.note_midi, d.note_dur, d.note_rest = nm_t.data_ptr(), nd_t.data_ptr(), nr_t.data_ptr()
        d.note_count = note_count.data_ptr()
        if dbg is not None:
            dbg['frame2item'] = torch.zeros(m, dtype=torch.int32, device=self.device)
            dbg['values'] = torch.zeros(m, dtype=torch.float32, device=self.device)
            dbg['rest'] = torch.zeros(m, dtype=torch.uint8, device=self.device)
            d.dbg_frame2item, d.dbg_values, d.dbg_rest = (dbg[k].data_ptr() for k in ('frame2item', 'values', 'rest'))
        d.scratch = ws.scratch.data_ptr()
        with self._mark('some_decode_notes', m * (self.outdim * 4.0 + 4.0)):
            _lib.check(self.lib.some_decode_notes(C.byref(d), self._stream), 'some_decode_notes')

    # ------------------------------------------------------------------ public batched entry point
    def tables(self, lens: np.ndarray):
        """Var-len tables for clips of ``lens`` samples: 16-byte aligned starts, cu_frames (T = 1 + L // hop,
        spec.py:48-60).  Host only."""
        lens = np.asarray(lens, dtype=np.int64)
        padded = (lens + 3) & ~3
        starts = np.zeros(len(lens), dtype=np.int64)
        np.cumsum(padded[:-1], out=starts[1:])
        cu = np.zeros(len(lens) + 1, dtype=np.int32)
        np.cumsum(1 + lens // HOP, out=cu[1:])
        return starts, lens, cu, int(padded.sum())

    def pack(self, waveforms: Sequence[np.ndarray]):
        """Concatenates clips into one pinned host buffer (test / bench helper).  Returns
        (pinned wave f32, pinned [starts | lens] int64, cu_frames int32 numpy)."""
        starts, lens, cu, total = self.tables([int(w.shape[0]) for w in waveforms])
        host = torch.empty(max(total, 4), dtype=torch.float32).pin_memory()
        for s, w, n in zip(starts, waveforms, lens):
            host[s:s + n].copy_(torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)))
        tables = torch.from_numpy(np.concatenate([starts, lens])).pin_memory()
        return host, tables, cu

    def _staging(self, total: int, b: int, m: int):
        """Grow-only pinned staging + device input buffers (cudaHostAlloc per call would dominate the step)."""
        st = getattr(self, '_stage', None)
        if st is None or st['wave_h'].numel() < total or st['tab_h'].numel() < 4 * b + 8 or st['out_h'].numel() < 9 * m + 4 * b + 64:
            cap_w = max(total, 4, int(1.25 * st['wave_h'].numel()) if st else 0)
            cap_b = max(4 * b + 8, st['tab_h'].numel() if st else 0)
            cap_o = max(9 * m + 4 * b + 64, int(1.25 * st['out_h'].numel()) if st else 0)
            st = {
                'wave_h': torch.empty(cap_w, dtype=torch.float32).pin_memory(),
                'wave_d': torch.empty(cap_w, dtype=torch.float32, device=self.device),
                'tab_h': torch.empty(cap_b, dtype=torch.int64).pin_memory(),
                'tab_d': torch.empty(cap_b, dtype=torch.int64, device=self.device),
                'out_h': torch.empty(cap_o, dtype=torch.uint8).pin_memory(),
                'out_d': torch.empty(cap_o, dtype=torch.uint8, device=self.device),
            }
            self._stage = st
        return st

    # Pipeline chunks of a large batch: staging + H2D of chunk c+1 overlap the kernels of chunk c.  Small chunks cost
    # kernel efficiency (measured on 64 x 30 s: 1 / 2 / 4 / 8 equal chunks -> 38.0 / 38.9 / 40.7 / 44.9 ms of kernels), so
    # the split is geometric: a small first chunk gets the GPU going, the later ones stay big.
    CHUNK_FRACTIONS = (0.125, 0.375, 0.5)
    MIN_CHUNK_FRAMES = 16384

    def _chunks(self, cu: np.ndarray) -> List[tuple]:
        b, m = len(cu) - 1, int(cu[-1])
        if b < 2 or m < 2 * self.MIN_CHUNK_FRAMES:
            return [(0, b)]
        bounds, acc = [0], 0.0
        for f in self.CHUNK_FRACTIONS[:-1]:
            acc += f
            i = int(np.searchsorted(cu, acc * m, side='left'))
            i = min(max(i, bounds[-1] + 1), b - 1)
            if cu[i] - cu[bounds[-1]] >= self.MIN_CHUNK_FRAMES // 2 and i > bounds[-1]:
                bounds.append(i)
        bounds.append(b)
        return [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1) if bounds[i + 1] > bounds[i]]

    def _pool(self):
        if getattr(self, '_tp', None) is None:
            import concurrent.futures
            import os
            # staging threads: memcpy-bound; leave cores to the other ranks of a one-process-per-GPU job
            local_world = int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1)
            cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 4)
            self._tp = concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(2, cores // (2 * local_world))))
        return self._tp

    def slab_layout(self, lens: Sequence[int]):
        """Deterministic layout of the packed note slab of a batch with these clip lengths: per pipeline chunk
        (c0, c1, byte offset, clips, frames) and the total size.  Chunk slab = [counts i32 [bc] | dur i32 [mc] | midi f32 [mc]
        | rest u8 [mc]].  Every rank of a data-parallel job can compute every other rank's layout from the lengths alone."""
        _, _, cu, _ = self.tables(lens)
        layout, off = [], 0
        for c0, c1 in self._chunks(cu):
            bc, mc = c1 - c0, int(cu[c1] - cu[c0])
            layout.append((c0, c1, off, bc, mc))
            off += (4 * bc + 9 * mc + 15) & ~15
        return cu, layout, off

    def enqueue(self, waveforms: Sequence[np.ndarray], quantized: bool = False, return_intermediates: bool = False,
                resident=None):
        """Stages, copies and enqueues the whole batch WITHOUT synchronising.  Returns (device slab uint8 [nbytes], cu, layout,
        extra): the decoded notes land in the device slab (see slab_layout); the caller copies it to the host (infer) or
        hands it to the all-gather (dist.infer_sharded).

        ``resident = (wave_d, ranges)``: the audio is already on the device (f32 tensor) and the clips are the sample ranges
        ``[(begin, end), ...]`` inside it (``waveforms`` is ignored): no staging, no audio H2D — the slicer path
        (infer_sliced) cuts the recording where it lies."""
        dev = self.device
        if resident is not None:
            wave_res, ranges = resident
            b = len(ranges)
            starts = np.asarray([r[0] for r in ranges], dtype=np.int64)
            lens = np.asarray([r[1] - r[0] for r in ranges], dtype=np.int64)
            cu = np.zeros(b + 1, dtype=np.int32)
            np.cumsum(1 + lens // HOP, out=cu[1:])
            total = 0
        else:
            b = len(waveforms)
            starts, lens, cu, total = self.tables([int(w.shape[0]) for w in waveforms])
        m = int(cu[-1])
        if return_intermediates:
            layout, nbytes_total = [(0, b, 0, b, m)], (4 * b + 9 * m + 15) & ~15
        else:
            _, layout, nbytes_total = self.slab_layout(lens)
        st = self._staging(total, b, m)
        stream = torch.cuda.current_stream(dev)
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
        copy_stream = self._copy_stream
        copy_stream.wait_stream(stream)       # previous users of the staging / device buffers are done
        wave_h, wave_d, tab_h, tab_d, out_d = (st[k] for k in ('wave_h', 'wave_d', 'tab_h', 'tab_d', 'out_d'))
        if resident is not None:
            wave_d = wave_res
            direct = []
        else:
            hv = wave_h.numpy()
            pool = self._pool()

            def stage(i):
                n = int(lens[i])
                if n:
                    hv[starts[i]:starts[i] + n] = waveforms[i]          # dtype cast (if any) + memcpy, GIL released

            # Clips that already live in page-locked memory (pinned_array(), or any float32 view of a pinned torch tensor)
            # are copied H2D straight from the caller's buffer: no staging memcpy at all.  Everything else goes through the
            # pinned staging buffer.
            direct = [bool(lens[i]) and w.dtype == np.float32 and w.flags.c_contiguous and torch.from_numpy(w).is_pinned()
                      for i, w in enumerate(waveforms)]

        ws = self.workspace(max(mc for *_, mc in layout))
        extra = None
        for c0, c1, out_off, bc, mc in layout:
            lo = int(starts[c0])
            todo = []
            if resident is None:
                todo = [i for i in range(c0, c1) if not direct[i]]
                if todo:
                    list(pool.map(stage, todo))
                hi = int(starts[c1 - 1] + ((lens[c1 - 1] + 3) & ~3))
            # var-len tables of this chunk, relative to its own first sample / first frame
            tab = tab_h[4 * c0:4 * c0 + 3 * bc + 1]
            tab[:bc] = torch.from_numpy(starts[c0:c1] - lo)
            tab[bc:2 * bc] = torch.from_numpy(lens[c0:c1])
            tab[2 * bc:3 * bc + 1] = torch.from_numpy((cu[c0:c1 + 1] - cu[c0]).astype(np.int64))
            tab_dev = tab_d[4 * c0:4 * c0 + 3 * bc + 1]
            with torch.cuda.stream(copy_stream):
                if resident is not None:
                    pass
                elif len(todo) == c1 - c0:
                    if hi > lo:
                        wave_d[lo:hi].copy_(wave_h[lo:hi], non_blocking=True)
                else:
                    for i in range(c0, c1):
                        n = int(lens[i])
                        if n:
                            src = torch.from_numpy(waveforms[i]) if direct[i] else wave_h[starts[i]:starts[i] + n]
                            wave_d[starts[i]:starts[i] + n].copy_(src, non_blocking=True)
                tab_dev.copy_(tab, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            stream.wait_event(ev)
            cu_d = tab_dev[2 * bc:3 * bc + 1].to(torch.int32)
            max_frames = int(np.diff(cu[c0:c1 + 1]).max())
            # decode writes straight into this chunk's slab
            o = out_d[out_off:out_off + 4 * bc + 9 * mc]
            note_count = o[:4 * bc].view(torch.int32)
            note_dur = o[4 * bc:4 * bc + 4 * mc].view(torch.int32)
            note_midi = o[4 * bc + 4 * mc:4 * bc + 8 * mc].view(torch.float32)
            note_rest = o[4 * bc + 8 * mc:]
            mel_f32 = torch.empty((mc, 80), dtype=torch.float32, device=dev) if return_intermediates else None
            self.run_mel(wave_d[lo:], tab_dev[:bc], tab_dev[bc:2 * bc], cu_d, bc, max_frames, mel_f32, ws.units)
            self.run_trunk(ws, mc, bc, cu_d, max_frames, 'softmax' if quantized else 'sigmoid')
            self.run_decode(ws, mc, bc, cu_d, note_count, quantized, out=(note_midi, note_dur, note_rest))
            if return_intermediates:
                extra = (mel_f32, ws.probs[:mc], ws.bounds[:mc])
        return out_d[:nbytes_total], cu, layout, extra

    def rms_frames(self, wave_d: torch.Tensor, frame_length: int, hop: int) -> np.ndarray:
        """Short-time RMS of a device-resident f32 waveform (some_slicer_rms), bit-identical to the reference's get_rms
        (utils/slicer2.py:5-38).  Returns the host copy (one small D2H + sync): the slicer's state machine runs on the host."""
        n = int(wave_d.numel())
        n_frames = 1 + (n + 2 * (frame_length // 2) - frame_length) // hop
        rms_d = torch.empty(n_frames, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.some_slicer_rms(wave_d.data_ptr(), n, frame_length, hop, rms_d.data_ptr(), n_frames, self._stream),
                   'some_slicer_rms')
        self.launches += 1
        return rms_d.cpu().numpy()

    def rms_frames_many(self, waves_d: Sequence[torch.Tensor], frame_length: int, hop: int) -> List[np.ndarray]:
        """RMS lists of several device-resident recordings: one launch each into ONE buffer, one D2H copy and sync for all
        (the dataset driver, some_b200/batch.py)."""
        counts = [1 + (int(w.numel()) + 2 * (frame_length // 2) - frame_length) // hop for w in waves_d]
        if not counts:
            return []
        cuts = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=cuts[1:])
        rms_d = torch.empty(int(cuts[-1]), dtype=torch.float32, device=self.device)
        for w, a, n in zip(waves_d, cuts, counts):
            _lib.check(self.lib.some_slicer_rms(w.data_ptr(), int(w.numel()), frame_length, hop, rms_d[a:].data_ptr(), n,
                                                self._stream), 'some_slicer_rms')
            self.launches += 1
        host = rms_d.cpu().numpy()
        return [host[a:b] for a, b in zip(cuts[:-1], cuts[1:])]

    def infer_sliced(self, waveform: np.ndarray, slicer, quantized: bool = False):
        """One long mono recording -> (chunk offsets in seconds, per-chunk notes): the flow of infer.py:38-41 /
        batch_infer.py:50-54 (Slicer.slice, then infer on the chunks) with the recording uploaded ONCE.  The RMS frames are
        computed on the device, the slicer's decisions are taken on the host from the copied RMS list (15 k floats for 5 min),
        and the chunks are then processed where they lie in device memory as one var-len batch."""
        from .slicer import chunk_ranges, silence_tags
        samples = np.ascontiguousarray(waveform, dtype=np.float32)
        assert samples.ndim == 1, 'infer_sliced expects a mono waveform (infer.py loads with mono=True)'
        n = int(samples.shape[0])
        dev = self.device
        with torch.cuda.device(dev):
            if (n + slicer.hop_size - 1) // slicer.hop_size <= slicer.min_length:   # slicer2.py:79-80
                return [0], self.infer([samples], quantized)
            src = torch.from_numpy(samples)
            if not src.is_pinned():
                st = self._staging(n, 1, 1 + n // HOP)
                st['wave_h'][:n].copy_(src)
                src = st['wave_h'][:n]
            wave_d = torch.empty(n, dtype=torch.float32, device=dev)
            wave_d.copy_(src, non_blocking=True)
            rms = self.rms_frames(wave_d, slicer.win_size, slicer.hop_size)
            ranges = chunk_ranges(silence_tags(rms, slicer), rms.shape[0], slicer.hop_size, n)
            if not ranges:
                return [], []
            slab, cu, layout, _ = self.enqueue(None, quantized, resident=(wave_d, ranges))
            self._staging(0, len(ranges), int(cu[-1]))
            out_h = self._stage['out_h']
            out_h[:slab.numel()].copy_(slab, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            results = self.unpack_slab(out_h[:slab.numel()].numpy(), cu, layout)
        offsets = [begin / slicer.sr for begin, _ in ranges]
        return offsets, results

    def unpack_slab(self, host: np.ndarray, cu: np.ndarray, layout, extra=None) -> List[Dict[str, np.ndarray]]:
        """Host copy of a packed note slab -> one dict per clip (input order).  Conversions are done once per chunk on the
        whole arrays; the per-clip entries are slices."""
        results: List[Dict[str, np.ndarray]] = []
        for c0, c1, off, bc, mc in layout:
            h = host[off:off + 4 * bc + 9 * mc]
            results.extend(self.unpack(cu[c0:c1 + 1] - cu[c0], h[:4 * bc].view(np.int32),
                                       h[4 * bc + 4 * mc:4 * bc + 8 * mc].view(np.float32),
                                       h[4 * bc:4 * bc + 4 * mc].view(np.int32), h[4 * bc + 8 * mc:4 * bc + 9 * mc], extra))
        return results

    def infer(self, waveforms: Sequence[np.ndarray], quantized: bool = False,
              return_intermediates: bool = False) -> List[Dict[str, np.ndarray]]:
        """waveform-in -> notes-out for a list of clips: the batched equivalent of BaseInference.infer
        (base_infer.py:46-53).  Host buffers in, host buffers out.  The batch is cut into up to 3 chunks of whole clips
        (small first chunk); for each chunk the clips are staged into pinned memory by a small thread pool (memcpy releases
        the GIL), copied H2D on a copy stream, and the kernels of the chunk are enqueued behind an event — so staging and
        H2D of chunk c+1 overlap the kernels of chunk c.  The notes come back in ONE packed D2H copy
        [counts | dur | midi | rest] per chunk slab; there is a single host synchronisation at the end."""
        if len(waveforms) == 0:
            return []
        dev = self.device
        with torch.cuda.device(dev):
            slab, cu, layout, extra = self.enqueue(waveforms, quantized, return_intermediates)
            out_h = self._stage['out_h']
            out_h[:slab.numel()].copy_(slab, non_blocking=True)
            if extra is not None:
                extra = tuple(t.cpu() for t in extra)
            torch.cuda.current_stream(dev).synchronize()
            return self.unpack_slab(out_h[:slab.numel()].numpy(), cu, layout, extra)

    def unpack(self, cu, nc, nm, nd, nr, extra=None) -> List[Dict[str, np.ndarray]]:
        dur_s = nd * self.timestep     # me_infer.py:95: int64 * python float -> float64; int32 * float gives the same float64s
        rest_b = nr.astype(bool)
        midi = nm.copy()                                     # the pinned staging buffer is reused by the next call
        first, count = cu[:-1].tolist(), nc.tolist()
        if extra is None:
            return [{'note_midi': midi[r0:r0 + n], 'note_dur': dur_s[r0:r0 + n], 'note_rest': rest_b[r0:r0 + n]}
                    for r0, n in zip(first, count)]
        out = []
        for i, (r0, n) in enumerate(zip(first, count)):
            e1 = int(cu[i + 1])
            out.append({'note_midi': midi[r0:r0 + n], 'note_dur': dur_s[r0:r0 + n], 'note_rest': rest_b[r0:r0 + n],
                        'mel': extra[0][r0:e1].numpy(), 'probs': extra[1][r0:e1].numpy(), 'bounds': extra[2][r0:e1].numpy()})
        return out
