"""Seeded stand-ins for decoded notes and a synthetic DiffSinger-style dataset (shared by make_golden_host_rows.py and the
tests of some_b200/midi.py and some_b200/batch.py)."""
import numpy as np

SEGMENT_SEEDS = (1, 2, 3, 4, 5)
TIMESTEP = 512 / 44100


def fake_segments(seed: int):
    """(chunk offsets in seconds, [{'note_midi' f32, 'note_dur' f64 = frames * timestep, 'note_rest' bool}]) shaped like
    MIDIExtractionInference.infer output (me_infer.py:90-97) for a sliced recording."""
    rng = np.random.default_rng(1000 + seed)
    n_chunks = int(rng.integers(1, 6))
    offsets, segments, t = [], [], float(rng.uniform(0.0, 0.4)) if seed % 2 else 0.0
    for c in range(n_chunks):
        n = int(rng.integers(0 if seed == 4 else 1, 14))
        frames = rng.integers(1, 90, size=n)
        dur = frames.astype(np.int64) * TIMESTEP
        midi = rng.uniform(45, 80, size=n).astype(np.float32)
        if seed == 3:
            midi = np.round(midi)                                   # quantised model: integer-valued float32
        rest = rng.random(n) < 0.2
        offsets.append(round(t * 50) * 882 / 44100)                 # slicer offsets are multiples of the 20 ms hop
        segments.append({'note_midi': midi, 'note_dur': dur, 'note_rest': rest})
        gap = float(rng.uniform(-0.3, 1.2))                         # negative: the next chunk starts before this one's notes end
        t = t + float(dur.sum()) + gap
    return np.asarray(offsets, dtype=np.float64), segments


def make_dataset_rows():
    """transcriptions.csv rows: name, phoneme sequence / durations / per-word phoneme counts."""
    rows = []
    for i, seed in enumerate(SEGMENT_SEEDS):
        rng = np.random.default_rng(2000 + seed)
        offsets, segments = fake_segments(seed)
        total = float(offsets[-1] + segments[-1]['note_dur'].sum()) if len(offsets) else 1.0
        n_words = int(rng.integers(3, 12))
        ph_num = rng.integers(1, 4, size=n_words)
        n_ph = int(ph_num.sum())
        cuts = np.sort(rng.uniform(0, max(total, 0.5), size=n_ph - 1))
        durs = np.diff(np.concatenate([[0.0], cuts, [max(total, 0.5)]]))
        rows.append({'name': f'item{i}', 'seed': seed, 'ph_seq': ' '.join(['a'] * n_ph),
                     'ph_dur': ' '.join(f'{d:.6f}' for d in durs), 'ph_num': ' '.join(str(int(x)) for x in ph_num)})
    rows.insert(2, {'name': 'missing', 'seed': 0, 'ph_seq': 'a', 'ph_dur': '0.5', 'ph_num': '1'})
    return rows
