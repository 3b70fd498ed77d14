"""Where does the end-to-end time of Engine.infer go?  (host staging / H2D / kernels)  Run on the GPU box."""
import sys, time, tempfile
import numpy as np, torch
sys.path.insert(0, '.')
from some_b200 import synth, plugin
import bench
cfg = synth.named_config('two_head')
with tempfile.TemporaryDirectory() as d:
    ins = plugin.MIDIExtractionInference(config=cfg, model_path=synth.write_checkpoint(d, cfg), device='cuda:0')
eng = ins.model
clips = bench.make_clips(0, 64, 30.0)
def t(f, n=4):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for nchunk in (1, 2, 4, 8):
    eng.MAX_CHUNKS = nchunk
    eng.MIN_CHUNK_FRAMES = 1024
    tot = t(lambda: eng.infer(clips))
    # GPU-only: same chunk sizes, data resident, no staging
    per = 64 // nchunk
    host, tables, cu = eng.pack(clips[:per])
    b, m, mf = per, int(cu[-1]), int(np.diff(cu).max())
    wave, td, cud = host.cuda(), tables.cuda(), torch.from_numpy(cu).cuda()
    ws = eng.workspace(m); nc = torch.empty(b, dtype=torch.int32, device='cuda')
    def dev():
        for _ in range(nchunk):
            eng.run_mel(wave, td[:b], td[b:], cud, b, mf, None, ws.units)
            eng.run_trunk(ws, m, b, cud, mf, 'sigmoid')
            eng.run_decode(ws, m, b, cud, nc, False)
    g = t(dev)
    t0 = time.perf_counter()
    for _ in range(10): dev()
    cpu_issue = (time.perf_counter() - t0) / 10 * 1e3
    torch.cuda.synchronize()
    print(f'chunks={nchunk}: infer {tot:.1f} ms | kernels only {g:.1f} ms | cpu issue time of the launches {cpu_issue:.1f} ms')
