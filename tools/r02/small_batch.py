"""Round-2 measurement (GPU box): single-clip / small-batch latency through plugin.infer() with and without CUDA-graph replay
(C1: one 10 s clip; also 1 x 30 s and 4 x 10 s): end to end (wall) and device-timed (CUDA events around the resident kernels)."""
import os, sys, time, tempfile, pathlib, json
import numpy as np, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[2]))
from some_b200 import synth, plugin
from some_b200.engine import pinned_array

config = synth.named_config('two_head')
with tempfile.TemporaryDirectory() as d:
    ckpt = synth.write_checkpoint(d, config, seed=1234)
    ins = plugin.MIDIExtractionInference(config=config, model_path=ckpt)
eng = ins.model
dev = eng.device
def pin(w):
    a = pinned_array(len(w)); a[:] = w; return a
cases = {'1x10s': [pin(synth.synth_waveform(11, seconds=10.0))], '1x30s': [pin(synth.synth_waveform(12, seconds=30.0))],
         '4x10s': [pin(synth.synth_waveform(20 + i, seconds=10.0)) for i in range(4)]}
out = {}
MODES = [(False, False, 'off'), (True, False, 'off'), (True, False, 'small'), (True, True, 'off'), (True, True, 'small')]
for name, clips in cases.items():
    secs = sum(len(c) for c in clips) / synth.SR
    for graphs, fold, pdl in MODES:
        eng.use_graphs = graphs
        eng.pdl = pdl
        if eng.ln_fold != fold:
            eng.set_ln_fold(fold)
        ref = ins.infer(clips)
        for _ in range(4):
            res = ins.infer(clips)
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            t0 = time.perf_counter(); res = ins.infer(clips); ts.append(time.perf_counter() - t0)
        ms = 1e3 * float(np.median(ts))
        same = all(np.array_equal(a[k], b[k]) for a, b in zip(ref, res) for k in ('note_midi', 'note_dur', 'note_rest'))
        # device-timed: audio resident, the kernels of one step (mel -> trunk -> decode), eager launches vs one graph replay
        host, tables, cu = eng.pack(clips)
        b, m, mf = len(clips), int(cu[-1]), int(np.diff(cu).max())
        wave, tab, cu_d = host.to(dev), tables.to(dev), torch.from_numpy(cu).to(dev)
        ws = eng.workspace(m); nc = torch.empty(b, dtype=torch.int32, device=dev)
        def step():
            eng.run_mel(wave, tab[:b], tab[b:], cu_d, b, mf, None, ws.units)
            was = eng.lib.some_set_pdl(1 if pdl != 'off' else 0)
            eng.run_trunk(ws, m, b, cu_d, mf, 'sigmoid')
            eng.lib.some_set_pdl(was)
            eng.run_decode(ws, m, b, cu_d, nc, False)
        for _ in range(3): step()
        torch.cuda.synchronize()
        if graphs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g): step()
            run = g.replay
        else:
            run = step
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        dms = e0.elapsed_time(e1) / 20
        out[f'{name} graphs={int(graphs)} fold={int(fold)} pdl={pdl}'] = {'e2e_ms': round(ms, 3), 'e2e_rtf': round(secs / (ms / 1e3)), 'device_ms': round(dms, 3),
                                               'device_rtf': round(secs / (dms / 1e3)), 'same_notes': same}
        print(name, 'graphs', graphs, 'fold', fold, 'pdl', pdl, f'e2e {ms:.3f} ms ({secs / (ms / 1e3):.0f} x RT)  device {dms:.3f} ms ({secs / (dms / 1e3):.0f} x RT)  same={same}')
print(json.dumps(out))
