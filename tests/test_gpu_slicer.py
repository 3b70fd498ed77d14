"""GPU tests of the slicer row: some_slicer_rms is BIT-identical to numpy's get_rms, the device slicer
cuts the golden chunks of the unmodified reference, and Engine.infer_sliced (recording uploaded once, chunks processed in
place) returns exactly what infer() returns for the separately uploaded chunks."""
import pathlib
import sys

import numpy as np
import pytest
import torch

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / 'golden'))

from oracle import slicer as osl  # noqa: E402
from slicer_cases import CASES, make_case  # noqa: E402
from some_b200 import plugin, synth  # noqa: E402
from some_b200 import slicer as psl  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(HERE / 'golden' / 'slicer.npz')


@pytest.mark.parametrize('n,frame,hop', [(441000, 3528, 882), (100000, 1024, 256), (5000, 3528, 882), (70000, 441, 441),
                                         (12345, 7, 3), (40000, 129, 64), (40000, 130, 50), (300, 2048, 512), (1, 8, 4)])
def test_rms_kernel_bit_exact(n, frame, hop):
    rng = np.random.default_rng(n + frame)
    x = (0.2 * rng.standard_normal(n)).astype(np.float32)
    x[n // 3:n // 3 + n // 10] = 0.0
    s = psl.Slicer(44100, max_sil_kept=1000)
    s.win_size, s.hop_size = frame, hop
    got = s.rms(x)
    ref = osl.rms_frames(x, frame, hop)
    assert got.dtype == np.float32 and got.shape == ref.shape
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize('name', list(CASES))
def test_device_slicer_cuts_the_reference_chunks(name):
    wave = make_case(name)
    s = psl.Slicer(44100, **CASES[name]['slicer'])
    chunks = s.slice(wave)
    got = np.asarray([(int(round(c['offset'] * 44100)), int(round(c['offset'] * 44100)) + len(c['waveform'])) for c in chunks],
                     dtype=np.int64).reshape(-1, 2)
    np.testing.assert_array_equal(got, GOLD[f'{name}__ranges'])
    np.testing.assert_array_equal(np.asarray([c['offset'] for c in chunks], dtype=np.float64), GOLD[f'{name}__offsets'])


@pytest.mark.parametrize('name', ['rec90', 'lead_trail', 'all_silence', 'short3'])
def test_infer_sliced_equals_infer_on_chunks(name, tmp_path):
    config = synth.named_config('two_head')
    ckpt = synth.write_checkpoint(tmp_path, config)
    ins = plugin.MIDIExtractionInference(config=config, model_path=ckpt, device='cuda:0')
    wave = make_case(name)
    s = psl.Slicer(44100, max_sil_kept=1000)
    offsets, notes = ins.model.infer_sliced(wave, s)
    ranges = GOLD[f'{name}__ranges']
    assert len(offsets) == len(ranges) == len(notes)
    np.testing.assert_array_equal(np.asarray(offsets, dtype=np.float64), GOLD[f'{name}__offsets'])
    if len(ranges):
        ref = ins.infer([wave[a:b] for a, b in ranges])
        for r, g in zip(ref, notes):
            for k in ('note_midi', 'note_dur', 'note_rest'):
                np.testing.assert_array_equal(r[k], g[k])
