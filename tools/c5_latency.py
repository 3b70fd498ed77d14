"""C5: one 5-minute recording -> slicer -> var-len batch of chunks -> notes; wall latency per recording.
Prints one JSON line.  Compares the device slicer path (Engine.infer_sliced) with the reference-style flow (numpy slicer on
the host, oracle/slicer.py == utils/slicer2.py, then plugin.infer on the chunks)."""
import json, pathlib, sys, tempfile, time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests' / 'golden'))
from some_b200 import plugin, synth  # noqa: E402
from some_b200 import slicer as psl  # noqa: E402
from some_b200.engine import pinned_array  # noqa: E402


def main():
    import contextlib
    from oracle import slicer as osl
    from slicer_cases import make_case
    config = synth.named_config('two_head')
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(sys.stderr):
        ckpt = synth.write_checkpoint(d, config)
        ins = plugin.MIDIExtractionInference(config=config, model_path=ckpt, device='cuda:0')
    wave = make_case('rec300')
    pw = pinned_array(len(wave))
    pw[:] = wave
    s = psl.Slicer(44100, max_sil_kept=1000)
    p = osl.SlicerParams(44100, max_sil_kept=1000)

    def timed(fn, n=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        t = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            t.append(time.perf_counter() - t0)
        return 1e3 * float(np.median(t))

    ms_sliced = timed(lambda: ins.model.infer_sliced(pw, s))
    ms_sliced_pageable = timed(lambda: ins.model.infer_sliced(wave, s))
    ranges = osl.slice_ranges(wave, p)
    ms_host_slicer = timed(lambda: osl.slice_ranges(wave, p), n=3, warm=1)
    ms_infer_chunks = timed(lambda: ins.infer([wave[a:b] for a, b in ranges]))
    ms_rms = timed(lambda: s.rms(wave))
    print(json.dumps({'workload': 'C5: two_head, one 300 s recording with silence gaps', 'chunks': len(ranges),
                      'infer_sliced_ms_pinned': ms_sliced, 'infer_sliced_ms_pageable': ms_sliced_pageable,
                      'reference_style_ms': ms_host_slicer + ms_infer_chunks, 'host_numpy_slicer_ms': ms_host_slicer,
                      'infer_on_chunks_ms': ms_infer_chunks, 'device_rms_incl_upload_ms': ms_rms,
                      'audio_seconds': len(wave) / 44100}))


if __name__ == '__main__':
    main()
