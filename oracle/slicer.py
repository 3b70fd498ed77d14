"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the reference's silence slicer,
utils/slicer2.py — ``get_rms`` (:5-38) and ``Slicer`` (:41-145) — used to check some_b200/slicer.py (GPU RMS + run-based
state machine).  Pinned to the unmodified reference by tests/golden/slicer.npz (tests/golden/make_golden_slicer.py).

Semantics kept exactly:
* rms: zero padding of frame_length // 2 on both sides, one frame every hop samples, sqrt(mean(|x|^2)) in float32 with
  numpy's reduction over the contiguous window axis (slicer2.py:13-38; the reference reduces axis -2 of an array whose
  axis -2 is the contiguous one, which is the same pairwise summation).
* frame walk (slicer2.py:84-127): a silence is a maximal run of frames with rms < threshold; when it ends at frame i it is
  cut if it is a long enough leading silence or (long enough and the current clip is long enough); the cut position(s) are
  first-index argmins of rms over windows bounded by max_sil_kept; trailing silence handled at :129-133.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def rms_frames(samples: np.ndarray, frame_length: int, hop_length: int) -> np.ndarray:
    half = int(frame_length // 2)
    padded = np.pad(samples, (half, half), mode='constant')
    windows = np.lib.stride_tricks.sliding_window_view(padded, frame_length)[::hop_length]
    return np.sqrt(np.mean(np.abs(windows) ** 2, axis=-1))


class SlicerParams:
    """The derived integer parameters of Slicer.__init__ (slicer2.py:42-60)."""

    def __init__(self, sr: int, threshold: float = -40., min_length: int = 5000, min_interval: int = 300,
                 hop_size: int = 20, max_sil_kept: int = 5000):
        if not min_length >= min_interval >= hop_size:
            raise ValueError('The following condition must be satisfied: min_length >= min_interval >= hop_size')
        if not max_sil_kept >= hop_size:
            raise ValueError('The following condition must be satisfied: max_sil_kept >= hop_size')
        interval_samples = sr * min_interval / 1000
        self.sr = sr
        self.threshold = 10 ** (threshold / 20.)
        self.hop_size = round(sr * hop_size / 1000)
        self.win_size = min(round(interval_samples), 4 * self.hop_size)
        self.min_length = round(sr * min_length / 1000 / self.hop_size)
        self.min_interval = round(interval_samples / self.hop_size)
        self.max_sil_kept = round(sr * max_sil_kept / 1000 / self.hop_size)


def silence_tags(rms: np.ndarray, p: SlicerParams) -> List[Tuple[int, int]]:
    """Frame ranges to remove, in order (slicer2.py:82-133), by the reference's frame-by-frame walk."""
    tags: List[Tuple[int, int]] = []
    sil_from = None
    clip_from = 0
    keep = p.max_sil_kept
    for i in range(rms.shape[0]):
        if rms[i] < p.threshold:
            if sil_from is None:
                sil_from = i
            continue
        if sil_from is None:
            continue
        leading = sil_from == 0 and i > keep
        middle = i - sil_from >= p.min_interval and i - clip_from >= p.min_length
        if leading or middle:
            span = i - sil_from
            if span <= keep:
                cut = int(rms[sil_from:i + 1].argmin()) + sil_from
                tags.append((0, cut) if sil_from == 0 else (cut, cut))
                clip_from = cut
            elif span <= 2 * keep:
                mid = int(rms[i - keep:sil_from + keep + 1].argmin()) + i - keep
                left = int(rms[sil_from:sil_from + keep + 1].argmin()) + sil_from
                right = int(rms[i - keep:i + 1].argmin()) + i - keep
                if sil_from == 0:
                    tags.append((0, right))
                    clip_from = right
                else:
                    tags.append((min(left, mid), max(right, mid)))
                    clip_from = max(right, mid)
            else:
                left = int(rms[sil_from:sil_from + keep + 1].argmin()) + sil_from
                right = int(rms[i - keep:i + 1].argmin()) + i - keep
                tags.append((0, right) if sil_from == 0 else (left, right))
                clip_from = right
        sil_from = None
    total = rms.shape[0]
    if sil_from is not None and total - sil_from >= p.min_interval:
        last = min(total, sil_from + keep)
        cut = int(rms[sil_from:last + 1].argmin()) + sil_from
        tags.append((cut, total + 1))
    return tags


def chunk_ranges(tags: List[Tuple[int, int]], total_frames: int, hop_size: int, num_samples: int) -> List[Tuple[int, int]]:
    """[begin, end) sample ranges kept by the tags (slicer2.py:62-71,135-145)."""
    if not tags:
        return [(0, num_samples)]
    frames = []
    if tags[0][0] > 0:
        frames.append((0, tags[0][0]))
    for a, b in zip(tags[:-1], tags[1:]):
        frames.append((a[1], b[0]))
    if tags[-1][1] < total_frames:
        frames.append((tags[-1][1], total_frames))
    return [(f0 * hop_size, min(num_samples, f1 * hop_size)) for f0, f1 in frames]


def slice_ranges(samples: np.ndarray, p: SlicerParams) -> List[Tuple[int, int]]:
    """Sample ranges of the chunks Slicer.slice would return for a mono waveform (slicer2.py:74-145)."""
    n = int(samples.shape[0])
    if (n + p.hop_size - 1) // p.hop_size <= p.min_length:
        return [(0, n)]
    rms = rms_frames(samples, p.win_size, p.hop_size)
    return chunk_ranges(silence_tags(rms, p), rms.shape[0], p.hop_size, n)
