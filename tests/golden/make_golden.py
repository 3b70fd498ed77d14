"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (/root/reference) on CPU
through oracle/refshim.py.  Runs only in the build container (the reference is not on the GPU
box); the outputs are committed so that tests can pin oracle/ (and the CUDA path) anywhere.

    python tests/golden/make_golden.py

Weights: some_b200.synth.fabricate_state_dict(config, seed=1234) loaded into the reference's own
``midi_conforms`` via load_state_dict(strict=True).  Waveforms: some_b200.synth (seeded).
Files are float32 / int64 / bool exactly as the reference returned them.
"""
import os
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


def main():
    torch.set_num_threads(1)  # deterministic reduction order for the committed numbers
    ref_inference = refshim.load_reference()
    import utils.infer_utils as ref_decode  # reference module (sys.path now has /root/reference first)
    import modules.rmvpe as ref_rmvpe

    # ---------------------------------------------------------------- decode known answers
    frame2item = torch.LongTensor([[1, 1, 1, 1, 2, 2, 3, 3, 3, 0, 0, 0, 0, 0],
                                   [1, 1, 1, 2, 3, 3, 3, 3, 3, 4, 4, 0, 0, 0]])
    values = torch.FloatTensor([[60, 61, 60.5, 63, 57, 57, 50, 55, 54, 0, 0, 0, 0, 0],
                                [50, 51, 50.5, 53, 47, 47, 40, 45, 44, 38, 38, 0, 0, 0]])
    iv, idur, im = ref_decode.decode_note_sequence(frame2item, values, frame2item > 0)
    bounds = torch.tensor([[.9, .05, .05, .6, .1, .3, .5, .02, .98, 0]])
    f2i = ref_decode.decode_bounds_to_alignment(bounds)
    probs = torch.zeros(1, 3, 128)
    probs[0, 0, 60], probs[0, 0, 61] = 0.9, 0.45
    probs[0, 1, 0], probs[0, 1, 3], probs[0, 1, 4] = 0.5, 0.5, 0.9
    probs[0, 2, 127] = 0.05
    gv, gr = ref_decode.decode_gaussian_blurred_probs(probs, vmin=0, vmax=127, deviation=1.0, threshold=0.1)
    # randomised decode cases: random probs / bounds straight into the reference decode functions
    g = torch.Generator().manual_seed(99)
    rb = torch.rand(4, 700, generator=g) ** 3
    rp = torch.rand(4, 700, 128, generator=g) ** 6
    rf2i = ref_decode.decode_bounds_to_alignment(rb)
    rv, rr = ref_decode.decode_gaussian_blurred_probs(rp, vmin=0, vmax=127, deviation=1.0, threshold=0.1)
    notes = [ref_decode.decode_note_sequence(rf2i[i:i + 1], rv[i:i + 1], ~rr[i:i + 1]) for i in range(4)]
    np.savez_compressed(
        HERE / 'decode_kat.npz',
        ex_frame2item=frame2item.numpy(), ex_values=values.numpy(),
        ex_item_values=iv.numpy(), ex_item_dur=idur.numpy(), ex_item_masks=im.numpy(),
        kat_bounds=bounds.numpy(), kat_frame2item=f2i.numpy(),
        kat_probs=probs.numpy(), kat_values=gv.numpy(), kat_rest=gr.numpy(),
        # inputs are regenerated in the tests from torch.Generator().manual_seed(99) (see above)
        rnd_bounds_checksum=np.float64(rb.double().sum().item()), rnd_probs_checksum=np.float64(rp.double().sum().item()),
        rnd_frame2item=rf2i.numpy(), rnd_values=rv.numpy(), rnd_rest=rr.numpy(),
        **{f'rnd_note_{k}_{i}': notes[i][j].numpy()[0] for i in range(4)
           for j, k in enumerate(('midi', 'dur', 'mask'))})
    print('decode_kat.npz written')

    # ---------------------------------------------------------------- mel front end
    mel_mod = ref_rmvpe.MelSpectrogram(n_mel_channels=80, sampling_rate=44100, win_length=2048,
                                       hop_length=512, mel_fmin=40, mel_fmax=8000)
    mel_out = {}
    clips = synth.edge_case_waveforms()
    clips['sung3s'] = synth.synth_waveform(101, seconds=3.0)
    for name, w in clips.items():
        with torch.no_grad():
            mel_out['mel_' + name] = mel_mod(torch.from_numpy(w).unsqueeze(0))[0].numpy()   # [80, T]
    np.savez_compressed(HERE / 'mel.npz', mel_basis=mel_mod.mel_basis.numpy(), **mel_out)
    print('mel.npz written')

    # ---------------------------------------------------------------- full plugin, per config
    for cfg_name, lay, secs, seeds in (('two_head', 3, 3.0, (201, 202)),
                                       ('quant_two_head', 3, 3.0, (203,)),
                                       ('midi_conformer', 8, 2.0, (204,))):
        config = synth.named_config(cfg_name)
        with tempfile.TemporaryDirectory() as d:
            ckpt = synth.write_checkpoint(d, config, seed=1234)
            cls = ref_inference.QuantizedMIDIExtractionInference if cfg_name.startswith('quant') \
                else ref_inference.MIDIExtractionInference
            ins = cls(config=synth.named_config(cfg_name), model_path=ckpt, device='cpu')
        out = {}
        waves = [synth.synth_waveform(s, seconds=secs + 0.37 * i) for i, s in enumerate(seeds)]
        if cfg_name == 'two_head':
            waves.append(synth.edge_case_waveforms()['ragged'])
            waves.append(synth.edge_case_waveforms()['short'])
        for i, w in enumerate(waves):
            sample = ins.preprocess(w)
            res = ins.forward_model(sample)
            out[f'clip{i}_probs'] = res['probs'][0].numpy().copy()
            out[f'clip{i}_bounds'] = res['bounds'][0].numpy().copy()
            notes_i = ins.postprocess(res)
            for k, v in notes_i.items():
                out[f'clip{i}_{k}'] = v
            out[f'clip{i}_num_samples'] = np.int64(len(w))
        # the public entry point must agree with the three-step form
        res_all = ins.infer(waves)
        for i, r in enumerate(res_all):
            for k, v in r.items():
                assert np.array_equal(v, out[f'clip{i}_{k}']), (cfg_name, i, k)
        out['seeds'] = np.array(seeds)
        out['seconds'] = np.float64(secs)
        np.savez_compressed(HERE / f'plugin_{cfg_name}.npz', **out)
        print(f'plugin_{cfg_name}.npz written', {k: v.shape for k, v in out.items() if 'note_midi' in k})


if __name__ == '__main__':
    main()
