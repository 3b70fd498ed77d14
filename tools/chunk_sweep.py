"""Developer sweep: end-to-end step time of plugin.infer() (C2 workload, pinned inputs) for several pipeline chunk splits."""
import contextlib, json, pathlib, sys, tempfile, time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from some_b200 import plugin, synth  # noqa: E402
from some_b200.engine import Engine, pinned_array  # noqa: E402


def main():
    config = synth.named_config('two_head')
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(sys.stderr):
        ckpt = synth.write_checkpoint(d, config)
        ins = plugin.MIDIExtractionInference(config=config, model_path=ckpt, device='cuda:0')
    clips = bench.make_clips(0, 64, 30.0)
    pinned = []
    for c in clips:
        a = pinned_array(len(c))
        a[:] = c
        pinned.append(a)
    out = {}
    for fr in [(0.125, 0.375, 0.5), (0.125, 0.875), (0.0625, 0.9375), (0.0625, 0.3125, 0.625), (0.03, 0.22, 0.75), (1.0,),
               (0.125, 0.375, 0.5)]:
        Engine.CHUNK_FRACTIONS = fr
        for src, name in ((pinned, 'pinned'), (clips, 'pageable')):
            for _ in range(2):
                ins.infer(src)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                ins.infer(src)
                ts.append(time.perf_counter() - t0)
            out[f'{fr} {name}'] = round(1e3 * float(np.median(ts)), 2)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
