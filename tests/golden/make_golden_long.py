"""Golden vectors AT THE BENCHMARKED SHAPES, produced by EXECUTING THE UNMODIFIED REFERENCE (/root/reference) on CPU
through oracle/refshim.py (build container only; the outputs are committed):

  * long_two_head.npz        one 30 s clip, configs/two_head_model.yaml         (BASELINE.json configs[1], T = 2584)
  * long_quant_two_head.npz  one 30 s clip, configs/quant_two_head_model.yaml   (configs[2])
  * long_midi_conformer.npz  one 10 s clip, configs/midi_conformer.yaml         (configs[3], lay 8, T = 862)
  * decode_quant_kat.npz     random 129-bin probabilities / bounds through the reference's quantised postprocess
                             (inference/me_quant_infer.py:21-38): bit-exact integer known answers for the decode kernel

The full probability matrices would be megabytes, so the long files hold: bounds [T] (complete), the decoded notes
(complete), per-frame max / argmax of probs and 96 complete rows of probs (evenly spaced).  tests/test_oracle_golden.py pins
oracle/ to them on the CPU; the GPU tests then compare the CUDA path with the (now long-clip pinned) oracle on every frame.

    python tests/golden/make_golden_long.py
"""
import pathlib
import sys
import tempfile

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))

from oracle import refshim  # noqa: E402
from some_b200 import synth  # noqa: E402

LONG_CASES = (('two_head', 30.0, 301), ('quant_two_head', 30.0, 302), ('midi_conformer', 10.0, 303))
N_ROWS = 96


def main():
    torch.set_num_threads(1)
    ref_inference = refshim.load_reference()
    for cfg_name, secs, seed in LONG_CASES:
        config = synth.named_config(cfg_name)
        with tempfile.TemporaryDirectory() as d:
            ckpt = synth.write_checkpoint(d, config, seed=1234)
            cls = ref_inference.QuantizedMIDIExtractionInference if cfg_name.startswith('quant') \
                else ref_inference.MIDIExtractionInference
            ins = cls(config=synth.named_config(cfg_name), model_path=ckpt, device='cpu')
        w = synth.synth_waveform(seed, seconds=secs)
        res = ins.forward_model(ins.preprocess(w))
        probs = res['probs'][0].numpy().copy()
        bounds = res['bounds'][0].numpy().copy()
        notes = ins.postprocess(res)
        t = probs.shape[0]
        rows = np.linspace(0, t - 1, N_ROWS).astype(np.int64)
        np.savez_compressed(HERE / f'long_{cfg_name}.npz', seed=np.int64(seed), seconds=np.float64(secs),
                            num_samples=np.int64(len(w)), bounds=bounds, probs_max=probs.max(1),
                            probs_argmax=probs.argmax(1).astype(np.int16), rows=rows, probs_rows=probs[rows],
                            **notes)
        print(f'long_{cfg_name}.npz: T={t} notes={len(notes["note_midi"])}')

    # quantised decode known answers: random inputs through the reference's own postprocess
    ins_q = None
    config = synth.named_config('quant_two_head')
    with tempfile.TemporaryDirectory() as d:
        ckpt = synth.write_checkpoint(d, config, seed=1234)
        ins_q = ref_inference.QuantizedMIDIExtractionInference(config=synth.named_config('quant_two_head'), model_path=ckpt,
                                                               device='cpu')
    g = torch.Generator().manual_seed(77)
    out = {}
    for i, t in enumerate((700, 700, 1, 37)):
        logits = torch.randn(1, t, 129, generator=g) * 2.0
        logits[..., 128] += 1.0                                    # a healthy share of rest frames
        probs = torch.softmax(logits, dim=-1)
        bounds = torch.rand(1, t, generator=g) ** 3
        res = ins_q.postprocess({'probs': probs.clone(), 'bounds': bounds.clone(), 'masks': torch.ones(1, t, dtype=torch.bool)})
        for k, v in res.items():
            out[f'q{i}_{k}'] = v
        out[f'q{i}_checksum'] = np.float64(probs.double().sum().item() + bounds.double().sum().item())
    np.savez_compressed(HERE / 'decode_quant_kat.npz', **out)
    print('decode_quant_kat.npz written', {k: v.shape for k, v in out.items() if 'midi' in k})


if __name__ == '__main__':
    main()
