"""CPU tests of the slicer row: the oracle restatement against the golden vectors produced by the
unmodified reference (tests/golden/make_golden_slicer.py), and the product's run-based host logic against the oracle's
frame-by-frame walk.  The RMS kernel itself is tested on the GPU (tests/test_gpu_slicer.py)."""
import pathlib
import sys

import numpy as np
import pytest

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / 'golden'))

from oracle import slicer as osl  # noqa: E402
from slicer_cases import CASES, make_case  # noqa: E402
from some_b200 import slicer as psl  # noqa: E402

GOLD = np.load(HERE / 'golden' / 'slicer.npz')


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_matches_reference_golden(name):
    wave = make_case(name)
    p = osl.SlicerParams(44100, **CASES[name]['slicer'])
    got = np.asarray(osl.slice_ranges(wave, p), dtype=np.int64).reshape(-1, 2)
    np.testing.assert_array_equal(got, GOLD[f'{name}__ranges'])
    np.testing.assert_array_equal(got[:, 0] / 44100, GOLD[f'{name}__offsets'])     # chunk['offset'] (slicer2.py:64)
    if f'{name}__rms' in GOLD:
        rms = osl.rms_frames(wave, p.win_size, p.hop_size)
        assert rms.shape[0] == int(GOLD[f'{name}__rms_len'])
        ref = GOLD[f'{name}__rms']
        np.testing.assert_array_equal(rms if ref.shape[0] == rms.shape[0] else rms[::7], ref)   # bit-exact
        assert float(rms.astype(np.float64).sum()) == float(GOLD[f'{name}__rms_sum'])


@pytest.mark.parametrize('name', [n for n in CASES if n not in ('short3',)])
def test_run_based_tags_equal_frame_walk_on_golden_cases(name):
    wave = make_case(name)
    kw = CASES[name]['slicer']
    p, s = osl.SlicerParams(44100, **kw), psl.Slicer(44100, **kw)
    rms = osl.rms_frames(wave, p.win_size, p.hop_size)
    assert psl.silence_tags(rms, s) == osl.silence_tags(rms, p)
    got = psl.chunk_ranges(psl.silence_tags(rms, s), rms.shape[0], s.hop_size, len(wave))
    np.testing.assert_array_equal(np.asarray(got, dtype=np.int64).reshape(-1, 2), GOLD[f'{name}__ranges'])


def test_run_based_tags_equal_frame_walk_random():
    """Random RMS lists with silent runs of every length class (<= keep, <= 2 keep, > 2 keep, leading, trailing, ties)."""
    rng = np.random.default_rng(4242)
    for trial in range(300):
        kw = dict(threshold=-40., min_length=int(rng.integers(300, 3000)), min_interval=int(rng.integers(40, 300)),
                  hop_size=int(rng.integers(5, 40)), max_sil_kept=int(rng.integers(40, 800)))
        if not kw['min_length'] >= kw['min_interval'] >= kw['hop_size'] or kw['max_sil_kept'] < kw['hop_size']:
            continue
        p, s = osl.SlicerParams(16000, **kw), psl.Slicer(16000, **kw)
        total = int(rng.integers(50, 1500))
        rms = (0.05 + 0.05 * rng.random(total)).astype(np.float32)
        pos = 0 if rng.random() < 0.3 else int(rng.integers(1, 40))
        while pos < total:
            ln = int(rng.integers(1, 120))
            quiet = (0.009 * rng.random(ln)).astype(np.float32)
            if rng.random() < 0.3:
                quiet[:] = np.float32(0.0)                                  # exact ties: first-index argmin must win
            rms[pos:pos + ln] = quiet[:max(0, min(ln, total - pos))]
            pos += ln + int(rng.integers(1, 200))
        assert psl.silence_tags(rms, s) == osl.silence_tags(rms, p), (trial, kw)


def test_slicer_parameters_and_errors():
    for kw in (dict(), dict(max_sil_kept=1000), dict(threshold=-30., min_length=2000, min_interval=200, hop_size=10, max_sil_kept=300)):
        p, s = osl.SlicerParams(44100, **kw), psl.Slicer(44100, **kw)
        for k in ('sr', 'threshold', 'hop_size', 'win_size', 'min_length', 'min_interval', 'max_sil_kept'):
            assert getattr(p, k) == getattr(s, k), k
    assert psl.Slicer(44100, max_sil_kept=1000).win_size == 3528 and psl.Slicer(44100).hop_size == 882
    with pytest.raises(ValueError):
        psl.Slicer(44100, min_length=100, min_interval=300)
    with pytest.raises(ValueError):
        psl.Slicer(44100, max_sil_kept=10)


def test_short_waveform_is_one_chunk_without_gpu():
    s = psl.Slicer(44100, max_sil_kept=1000)
    w = make_case('short3')
    chunks = s.slice(w)                  # below min_length: the reference returns the input untouched (slicer2.py:79-80)
    assert len(chunks) == 1 and chunks[0]['offset'] == 0 and chunks[0]['waveform'] is w
