"""CPU fp32 restatement of the SOME log-mel front end and two-head conformer forward, written
functionally over a plain ``state_dict`` (no nn.Module from the reference is used).

Each function cites the reference lines it follows (paths relative to /root/reference).
Pinned against the unmodified reference through tests/golden (see oracle/__init__.py).

TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import functools

import numpy as np
import torch
import torch.nn.functional as F

from .melbank import mel_filterbank


@functools.lru_cache(maxsize=8)
def _mel_basis(sr, n_fft, n_mels, fmin, fmax):
    """The reference builds the filterbank once in MelSpectrogram.__init__ (spec.py:22-29), not per clip."""
    return torch.from_numpy(mel_filterbank(sr, n_fft, n_mels, fmin, fmax)).float()


# --------------------------------------------------------------------------- mel front end
def log_mel(audio: torch.Tensor, sr=44100, n_fft=2048, hop=512, n_mels=80, fmin=40, fmax=8000,
            clamp=1e-5, keyshift=0, speed=1, center=True) -> torch.Tensor:
    """modules/rmvpe/spec.py:38-72.  audio [B, L] float32 -> log-mel [B, n_mels, T]; with the defaults (the inference path)
    T = 1 + L // hop.  keyshift / speed: the binarizer's augmentation path (spec.py:39-46,63-68)."""
    factor = 2 ** (keyshift / 12)                                      # spec.py:39
    n_fft_new = int(np.round(n_fft * factor))                          # :40 (win_length == n_fft in every shipped config)
    hop_new = int(np.round(hop * speed))                               # :42
    win = torch.hann_window(n_fft_new)                                 # :44-46 (periodic Hann)
    if center:
        audio = F.pad(audio, (n_fft_new // 2, (n_fft_new + 1) // 2))  # :47-50 zero pad
    fft = torch.stft(audio, n_fft=n_fft_new, hop_length=hop_new, win_length=n_fft_new, window=win,
                     center=False, return_complex=True)               # :52-60
    magnitude = fft.abs()                                             # :61
    if keyshift != 0:                                                  # :63-68
        size = n_fft // 2 + 1
        resize = magnitude.size(1)
        if resize < size:
            magnitude = F.pad(magnitude, (0, 0, 0, size - resize))
        magnitude = magnitude[:, :size, :] * n_fft / n_fft_new
    basis = _mel_basis(sr, n_fft, n_mels, fmin, fmax)                 # spec.py:22-29 (built once)
    mel_output = torch.matmul(basis, magnitude)                       # spec.py:70
    return torch.log(torch.clamp(mel_output, min=clamp))              # spec.py:71


# --------------------------------------------------------------------------- conformer trunk
def _lin(sd, name, x, bias=True):
    return F.linear(x, sd[name + '.weight'], sd[name + '.bias'] if bias else None)


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + '.weight'], sd[name + '.bias'], 1e-5)


def _glu(x, dim):
    out, gate = x.chunk(2, dim=dim)                                   # Gconform.py:15-18 / base_conv.py:12-15
    return out * gate.sigmoid()


def ffn(sd, p, x):
    """conform_ffn.forward, Gconform.py:29-34 (dropout = identity in eval)."""
    return _lin(sd, p + '.ln2', F.silu(_lin(sd, p + '.ln1', x)))


def attention(sd, p, x, heads=8):
    """Attention.forward, base_attention.py:23-46 with kv=None, mask=None (conform_blocke never
    passes a mask on the inference path: Gconform.py:83-84,133)."""
    b, t, _ = x.shape
    q = _lin(sd, p + '.to_q', x, bias=False)                          # :31
    k, v = _lin(sd, p + '.to_kv', x, bias=False).chunk(2, dim=2)     # :32  (k first)
    q, k, v = (z.reshape(b, t, heads, -1).transpose(1, 2) for z in (q, k, v))   # :34-36 'b t (h c) -> b h t c'
    out = F.scaled_dot_product_attention(q, k, v)                     # :41-43, scale = c ** -0.5
    out = out.transpose(1, 2).reshape(b, t, -1)                       # :45
    return _lin(sd, p + '.to_out.0', out)                             # :46


def conv_module(sd, p, x):
    """conform_conv.forward, base_conv.py:63-70; BatchNorm1d in eval mode (running stats)."""
    x = x.transpose(1, 2)                                             # :64
    x = _glu(F.conv1d(x, sd[p + '.pointwise_conv1.weight'], sd[p + '.pointwise_conv1.bias']), 1)  # :65
    k = sd[p + '.depthwise_conv.weight'].shape[-1]
    x = F.conv1d(x, sd[p + '.depthwise_conv.weight'], sd[p + '.depthwise_conv.bias'],
                 padding=(k - 1) // 2, groups=x.shape[1])             # :66
    x = F.batch_norm(x, sd[p + '.norm.running_mean'], sd[p + '.norm.running_var'],
                     sd[p + '.norm.weight'], sd[p + '.norm.bias'], False, 0.1, 1e-5)   # :67
    x = F.silu(x)                                                     # :68
    x = F.conv1d(x, sd[p + '.pointwise_conv2.weight'], sd[p + '.pointwise_conv2.bias'])  # :69
    return x.transpose(1, 2)                                          # :70


def conform_block(sd, p, x, heads=8, taps=None):
    """conform_blocke.forward, Gconform.py:56-63."""
    x = ffn(sd, p + '.ffn1', _ln(sd, p + '.norm1', x)) * 0.5 + x      # :57
    if taps is not None:
        taps[p + ':ffn1'] = x
    x = attention(sd, p + '.att', _ln(sd, p + '.norm2', x), heads) + x   # :60
    if taps is not None:
        taps[p + ':att'] = x
    x = conv_module(sd, p + '.conv', _ln(sd, p + '.norm3', x)) + x    # :61
    if taps is not None:
        taps[p + ':conv'] = x
    x = ffn(sd, p + '.ffn2', _ln(sd, p + '.norm4', x)) * 0.5 + x      # :62
    x = _ln(sd, p + '.norm5', x)                                      # :63
    if taps is not None:
        taps[p + ':out'] = x
    return x


def trunk(sd, units, lay, heads=8, mask=None, taps=None):
    """Gmidi_conform.forward, Gconform.py:119-140.  units [B, T, indim] -> (logits [B, T, outdim],
    bound probs [B, T]).  ``pitch`` is ignored by the reference and omitted here."""
    x = _lin(sd, 'model.inln', units)                                 # :124
    x1 = _lin(sd, 'model.inln1', units)                               # :122,125
    if mask is not None:
        x = x.masked_fill(~mask.unsqueeze(-1), 0)                     # :126-127
    for i in range(lay):                                              # :128-132, Gcf.forward :82-87
        p = f'model.cf_lay.{i}'
        midi = conform_block(sd, p + '.att1', x, heads, taps)
        bound = conform_block(sd, p + '.att2', x1, heads, taps)
        midis = _glu(_lin(sd, p + '.glu1.0', midi), 2)
        bounds = _glu(_lin(sd, p + '.glu2.0', bound), 2)
        x, x1 = midi + bounds, bound + midis
        if mask is not None:
            x = x.masked_fill(~mask.unsqueeze(-1), 0)
        if taps is not None:
            taps[p + ':midi'] = x
            taps[p + ':bound'] = x1
    x = conform_block(sd, 'model.att1', x, heads, taps)               # :133
    x1 = conform_block(sd, 'model.att2', x1, heads, taps)
    cut = torch.sigmoid(_lin(sd, 'model.cutheard', x1)).squeeze(-1)   # :135,137-138
    logits = _lin(sd, 'model.outln', x)                               # :136
    return logits, cut


def model_forward(sd, config, units, mask=None, sig=False, softmax=False, taps=None):
    """midi_conforms.forward, modules/model/Gmidi_conform.py:30-40."""
    args = config['midi_extractor_args']
    midi, bound = trunk(sd, units, args['lay'], args.get('attention_heads', 4), mask, taps)
    if taps is not None:
        taps['logits'] = midi
    if sig:
        midi = torch.sigmoid(midi)
    if softmax:
        midi = F.softmax(midi, dim=2)
    return midi, bound
