"""Restatement of ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=True)``
(librosa 0.9.x, default ``norm='slaney'``) — the one third-party arithmetic on the
hot path that is absent from /root/reference (requirements.txt:10 pins
``librosa<0.10.0``; call site modules/rmvpe/spec.py:22-28).

Published algorithm (librosa/filters.py::mel, librosa/core/convert.py):
  fftfreqs = linspace(0, sr/2, 1 + n_fft//2)
  mel_f    = mel_to_hz(linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2), htk=True)
  hz_to_mel(f) = 2595 * log10(1 + f / 700);  mel_to_hz(m) = 700 * (10**(m / 2595) - 1)
  fdiff = diff(mel_f); ramps = mel_f[:, None] - fftfreqs[None, :]
  w[i] = max(0, min(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
  slaney norm: w[i] *= 2 / (mel_f[i + 2] - mel_f[i])
ramps in float64, weights held in a float32 array (librosa's default dtype).

Parity: no reference test pins it; cross-checked in tests/test_oracle_golden.py against
torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='htk') (independent
implementation, agrees to ~1e-7) and the SURVEY.md §4 known answers (sum 3.716216,
727 non-zeros, bins 2..371).

TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import numpy as np


def hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    if fmax is None:
        fmax = float(sr) / 2
    fftfreqs = np.linspace(0.0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = mel_to_hz_htk(np.linspace(hz_to_mel_htk(fmin), hz_to_mel_htk(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    # librosa 0.9: the weight array is float32 from the start (dtype=np.float32 default),
    # ramps/fdiff/enorm are float64; each assignment / in-place multiply rounds to float32.
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights
