"""Import shim that makes the UNMODIFIED reference (/root/reference) importable in the
build container, where librosa / lightning / mido / parselmouth are not installed.

Used ONLY by tests/golden/make_golden.py (to generate committed fixtures) and by
container-only tests that pin oracle/ against the real reference.  /root/reference does
not exist on the GPU box, so nothing that runs there may call ``load_reference()``.

Stubs:
  * lightning(.pytorch{,.callbacks,.utilities{,.rank_zero}}): names only; the inference
    path never calls them (utils/__init__.py:11 -> utils/training_utils.py:7 imports them).
  * mido, parselmouth: names only (utils/infer_utils.py:3, utils/binarizer_utils.py:5).
  * librosa: ``filters.mel`` = oracle.melbank (restated librosa 0.9 formula); the rest are
    names only (modules/rmvpe/spec.py:4, inference/me_infer.py:4).

TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('SOME_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'inference', 'me_infer.py'))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    """Placeholder class usable as a base class, decorator or callable."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return self


def _passthrough_decorator(fn=None, *a, **k):
    if callable(fn):
        return fn
    return lambda f: f


def install_stubs():
    from . import melbank

    if 'lightning' not in sys.modules:
        rank_zero = _mod('lightning.pytorch.utilities.rank_zero',
                         rank_zero_only=_passthrough_decorator,
                         rank_zero_info=print, rank_zero_debug=lambda *a, **k: None,
                         rank_zero_warn=print)
        utilities = _mod('lightning.pytorch.utilities', rank_zero=rank_zero,
                         rank_zero_only=_passthrough_decorator)
        callbacks = _mod('lightning.pytorch.callbacks', ModelCheckpoint=_Anything,
                         TQDMProgressBar=_Anything)
        strategies = _mod('lightning.pytorch.strategies', DDPStrategy=_Anything)
        plm = _mod('lightning.pytorch', callbacks=callbacks, utilities=utilities,
                   strategies=strategies, LightningModule=_Anything, Trainer=_Anything,
                   seed_everything=lambda *a, **k: None)
        _mod('lightning', pytorch=plm)
    if 'mido' not in sys.modules:
        _mod('mido', MidiFile=_Anything, MidiTrack=_Anything, MetaMessage=_Anything,
             Message=_Anything, bpm2tempo=lambda bpm: int(round(60_000_000 / bpm)))
    if 'parselmouth' not in sys.modules:
        _mod('parselmouth', Sound=_Anything)
    if 'librosa' not in sys.modules:
        def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm='slaney', **kw):
            assert htk and norm == 'slaney', 'shim restates only the htk/slaney path the hot path uses'
            return melbank.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)

        filters = _mod('librosa.filters', mel=mel)
        sequence = _mod('librosa.sequence', viterbi=_Anything(), transition_local=_Anything())
        _mod('librosa', filters=filters, sequence=sequence, load=_Anything(),
             hz_to_midi=_Anything(), midi_to_hz=_Anything(), resample=_Anything())


def load_reference():
    """Returns the reference's ``inference`` package (unmodified code, CPU)."""
    if not reference_available():
        raise RuntimeError(f'reference not present at {REFERENCE_ROOT} (build container only)')
    install_stubs()
    # The reference must win over this repo's own drop-in ``inference`` package.
    for name in [n for n in sys.modules if n == 'inference' or n.startswith('inference.')]:
        del sys.modules[name]
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib
        ref_inference = importlib.import_module('inference')
        assert os.path.realpath(ref_inference.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), \
            f'wrong inference package imported: {ref_inference.__file__}'
        return ref_inference
    finally:
        pass
