"""fp32 VALIDATION mode (some_forward_f32, csrc/accurate.cu; SURVEY.md §7 hard part 3 (ii)): the trunk with fp32 operands on
the CUDA cores against the fp32 oracle — the "within 1e-3 fp32" line of the north-star contract — and the decoded notes, which
with fp32 arithmetic on both sides are expected to agree (almost) exactly: bf16 operand rounding is the only reason the
product path's notes differ from the reference's (tests/test_gpu_parity_long.py)."""
import numpy as np
import pytest
import torch

from some_b200 import synth
from oracle.metrics import note_agreement

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('cfg_name,seconds', [('two_head', 10.0), ('quant_two_head', 6.0), ('midi_conformer', 4.0)])
def test_fp32_mode_matches_oracle_to_1e_3(cfg_name, seconds, tmp_path):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from oracle import decode as odecode
    from some_b200 import plugin
    config = synth.named_config(cfg_name)
    ckpt = synth.write_checkpoint(tmp_path, config, seed=1234)
    quant = cfg_name.startswith('quant')
    cls = plugin.QuantizedMIDIExtractionInference if quant else plugin.MIDIExtractionInference
    ins = cls(config=config, model_path=ckpt)
    waves = [synth.synth_waveform(321, seconds=seconds), synth.synth_waveform(322, seconds=1.7)]
    got = ins.model.infer_accurate(waves, quantized=quant)
    sd = synth.fabricate_state_dict(config, seed=1234)
    for w, g in zip(waves, got):
        ref = odecode.infer_clip(sd, config, w, quantized=quant, return_intermediates=True)
        dp = float(np.abs(g['probs'] - ref['probs']).max())
        db = float(np.abs(g['bounds'] - ref['bounds']).max())
        fr, bd = note_agreement(ref, g)
        print(f'{cfg_name}: T={g["probs"].shape[0]} max|probs|={dp:.2e} max|bounds|={db:.2e} '
              f'mean(bounds err)={float((g["bounds"].astype(np.float64) - ref["bounds"]).mean()):+.2e} '
              f'frame agreement={fr:.4f} exact boundaries={bd:.4f} notes {len(g["note_midi"])} (oracle {len(ref["note_midi"])})')
        assert dp < 1e-3 and db < 1e-3, (cfg_name, dp, db)          # north_star: 1e-3 for fp32
        assert fr > 0.99 and bd > 0.97, (cfg_name, fr, bd)
