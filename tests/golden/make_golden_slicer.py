"""Generates tests/golden/slicer.npz by EXECUTING THE UNMODIFIED reference slicer (utils/slicer2.py) through
oracle/refshim.py (build container only).  Waveforms are regenerated in the tests from the same seeds.

    python tests/golden/make_golden_slicer.py
"""
import pathlib
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))

from oracle import refshim  # noqa: E402
sys.path.insert(0, str(HERE))
from slicer_cases import CASES, make_case  # noqa: E402


def main():
    refshim.load_reference()
    import utils.slicer2 as ref_slicer
    out = {}
    for name, spec in CASES.items():
        wave = make_case(name)
        sl = ref_slicer.Slicer(sr=44100, **spec['slicer'])
        chunks = sl.slice(wave)
        ranges = []
        for c in chunks:
            begin = int(round(c['offset'] * 44100))
            ranges.append((begin, begin + len(c['waveform'])))
            assert np.array_equal(wave[begin:begin + len(c['waveform'])], c['waveform'])
        out[f'{name}__ranges'] = np.asarray(ranges, dtype=np.int64).reshape(-1, 2)
        out[f'{name}__offsets'] = np.asarray([c['offset'] for c in chunks], dtype=np.float64)
        if (len(wave) + sl.hop_size - 1) // sl.hop_size > sl.min_length:
            rms = ref_slicer.get_rms(y=wave, frame_length=sl.win_size, hop_length=sl.hop_size).squeeze(0)
            out[f'{name}__rms_len'] = np.int64(rms.shape[0])
            # the whole list for small cases, a strided sample + checksum for the long ones
            out[f'{name}__rms'] = rms if rms.shape[0] <= 4000 else rms[::7]
            out[f'{name}__rms_sum'] = np.float64(rms.astype(np.float64).sum())
        print(name, len(wave), len(chunks), ranges[:4])
    np.savez_compressed(HERE / 'slicer.npz', **out)


if __name__ == '__main__':
    main()
