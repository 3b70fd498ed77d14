"""Agreement measures between two decoded note lists (test / bench reporting only).

TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import numpy as np


def note_agreement(a, b, hop=512 / 44100):
    """(frame-level agreement, exact-boundary agreement) of note lists ``a`` (reference) and ``b``: every frame carries the
    label of its note (rounded pitch, or rest); boundaries = the frames where a new note starts.  The decode is a
    discontinuous function of the probabilities (cumsum().round(), utils/infer_utils.py:28), so bf16-vs-fp32 parity of the
    decoded notes is a rate, not an equality."""
    def expand(r):
        d = np.rint(np.asarray(r['note_dur']) / hop).astype(int)
        lab = np.where(r['note_rest'], -1, np.rint(r['note_midi']).astype(int))
        return np.repeat(lab, d), np.cumsum(d)[:-1]
    ea, ba = expand(a)
    eb, bb = expand(b)
    n = min(len(ea), len(eb))
    frames = float((ea[:n] == eb[:n]).mean()) if n else 1.0
    bounds = float(np.isin(ba, bb).mean()) if len(ba) else 1.0
    return frames, bounds
