"""Seeded waveforms for the slicer parity tests (shared by tests/golden/make_golden_slicer.py and the tests)."""
import numpy as np

from some_b200 import synth

SR = 44100
INFER = dict(max_sil_kept=1000)                      # infer.py:39 / batch_infer.py:52
CASES = {
    'rec90': dict(slicer=INFER),                     # 90 s recording with 0.5-1.0 s digital-silence gaps
    'rec300': dict(slicer=INFER),                    # C5: 5 min
    'short3': dict(slicer=INFER),                    # shorter than min_length -> one chunk
    'lead_trail': dict(slicer=INFER),                # 3 s leading and 2.5 s trailing silence
    'long_gaps': dict(slicer=INFER),                 # gaps of 1.5 .. 4 s: the 2 * max_sil_kept and beyond branches
    'all_silence': dict(slicer=INFER),
    'no_silence': dict(slicer=INFER),
    'defaults': dict(slicer=dict()),                 # class defaults (max_sil_kept 5000)
    'tight': dict(slicer=dict(threshold=-30., min_length=2000, min_interval=200, hop_size=10, max_sil_kept=300)),
}


def _with_gaps(seed, seconds, gap_lo, gap_hi, every_lo, every_hi):
    rng = np.random.default_rng(seed)
    w = synth.synth_waveform(seed, seconds=seconds).copy()
    t = int(rng.uniform(every_lo, every_hi) * SR)
    while t < len(w):
        g = int(rng.uniform(gap_lo, gap_hi) * SR)
        w[t:t + g] = 0.0
        t += g + int(rng.uniform(every_lo, every_hi) * SR)
    return w


def make_case(name: str) -> np.ndarray:
    if name == 'rec90':
        return synth.synth_waveform(9001, seconds=90.0, silence_gaps=True)
    if name == 'rec300':
        return synth.synth_waveform(9002, seconds=300.0, silence_gaps=True)
    if name == 'short3':
        return synth.synth_waveform(9003, seconds=3.0)
    if name == 'lead_trail':
        w = synth.synth_waveform(9004, seconds=40.0, silence_gaps=True).copy()
        w[:3 * SR] = 0.0
        w[-int(2.5 * SR):] = 0.0
        return w
    if name == 'long_gaps':
        return _with_gaps(9005, 80.0, 1.5, 4.0, 5.5, 9.0)
    if name == 'all_silence':
        return np.zeros(20 * SR, dtype=np.float32)
    if name == 'no_silence':
        rng = np.random.default_rng(9006)
        return (0.1 * rng.standard_normal(30 * SR)).astype(np.float32)
    if name == 'defaults':
        return _with_gaps(9007, 70.0, 0.4, 7.0, 5.5, 8.0)
    if name == 'tight':
        return _with_gaps(9008, 45.0, 0.25, 1.0, 2.5, 4.0)
    raise KeyError(name)
