"""CPU tests of the host rows after the path against golden vectors produced by the unmodified
reference (tests/golden/make_golden_host_rows.py): MIDI message lists of build_midi_file and the CSV bytes of batch_infer.py."""
import csv
import io
import json
import pathlib
import struct
import sys

import numpy as np
import pytest

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / 'golden'))

from host_row_cases import SEGMENT_SEEDS, fake_segments, make_dataset_rows  # noqa: E402
from some_b200 import batch, midi  # noqa: E402

GOLD = np.load(HERE / 'golden' / 'host_rows.npz')


@pytest.mark.parametrize('seed', SEGMENT_SEEDS)
@pytest.mark.parametrize('tempo', [120, 97.5])
def test_midi_messages_equal_reference(seed, tempo):
    offsets, segments = fake_segments(seed)
    f = midi.build_midi_file(list(offsets), segments, tempo=tempo)
    mine = [(0, -1, d, a[0]) if k == 'set_tempo' else (1 if k == 'note_on' else 2, a[0], d, -1) for d, k, a in f.tracks[0]]
    np.testing.assert_array_equal(np.asarray(mine, dtype=np.int64).reshape(-1, 4), GOLD[f'midi_{seed}_{tempo}'])


def test_midi_wire_format_known_answer(tmp_path):
    """Hand-assembled SMF bytes (format 1, 480 tpb, set_tempo 500000, two notes, one clipped at the next chunk's offset)."""
    segs = [{'note_midi': np.array([60.2, 61.7, 55.0], np.float32), 'note_dur': np.array([0.5, 0.25, 0.3]),
             'note_rest': np.array([False, True, False])},
            {'note_midi': np.array([70.49], np.float32), 'note_dur': np.array([1.0]), 'note_rest': np.array([False])}]
    f = midi.build_midi_file([0.0, 1.0], segs, tempo=120)
    body = (b'\x00\xff\x51\x03\x07\xa1\x20' + b'\x00\x90\x3c\x40' + b'\x83\x60\x80\x3c\x40' + b'\x81\x70\x90\x37\x40'
            + b'\x81\x70\x80\x37\x40' + b'\x00\x90\x46\x40' + b'\x87\x40\x80\x46\x40' + b'\x00\xff\x2f\x00')
    want = b'MThd' + struct.pack('>IHHH', 6, 1, 1, 480) + b'MTrk' + struct.pack('>I', len(body)) + body
    assert f.to_bytes() == want
    f.save(tmp_path / 'a.mid')
    fmt, tpb, tracks = midi.parse_midi((tmp_path / 'a.mid').read_bytes())
    assert (fmt, tpb) == (1, 480) and [m[0] for m in tracks[0]] == [0, 0, 480, 240, 240, 0, 960, 0]
    assert midi._varlen(0) == b'\x00' and midi._varlen(127) == b'\x7f' and midi._varlen(128) == b'\x81\x00'
    assert midi._varlen(0x0fffffff) == b'\xff\xff\xff\x7f'
    with pytest.raises(ValueError):
        midi._varlen(-1)


def test_midi_roundtrip_random():
    for seed in SEGMENT_SEEDS:
        offsets, segments = fake_segments(seed)
        events = midi.note_events(list(offsets), segments, tempo=120)
        fmt, tpb, tracks = midi.parse_midi(midi.build_midi_file(list(offsets), segments).to_bytes())
        t, on, got = 0, {}, []
        for delta, status, data in tracks[0]:
            t += delta
            if status == 0x90:
                on[data[0]] = t
            elif status == 0x80:
                got.append((on.pop(data[0]), t, data[0]))
        assert got == events and all(a < b for a, b, _ in events)
        assert all(e0[1] <= e1[0] for e0, e1 in zip(events[:-1], events[1:]))          # monophonic, ordered


def test_calc_seq_equals_reference():
    for v, want in zip(GOLD['calc_seq_in'], GOLD['calc_seq_out']):
        assert batch.calc_seq(float(v), False) == str(want)
    assert batch.calc_seq(60.0, True) == 'rest' and batch.calc_seq(60.0, False) == 'C4' and batch.midi_to_note(0) == 'C-1'


@pytest.mark.parametrize('seed', SEGMENT_SEEDS)
def test_note_timeline_equals_reference(seed):
    offsets, segments = fake_segments(seed)
    assert batch.note_timeline(offsets, segments) == json.loads(str(GOLD['timelines_json']))[str(seed)]


@pytest.mark.parametrize('round_midi', [False, True])
def test_csv_bytes_equal_reference(round_midi):
    out = io.StringIO(newline='')
    writer = csv.DictWriter(out, fieldnames=['name', 'ph_seq', 'ph_dur', 'ph_num', 'note_seq', 'note_dur'])
    writer.writeheader()
    rows = []
    for r in make_dataset_rows():
        row = {k: r[k] for k in ('name', 'ph_seq', 'ph_dur', 'ph_num')}
        if r['name'] != 'missing':
            offsets, segments = fake_segments(r['seed'])
            row['note_seq'], row['note_dur'] = batch.row_notes(row['ph_dur'], row['ph_num'], batch.note_timeline(offsets, segments), round_midi)
        rows.append(row)
    writer.writerows(rows)
    assert out.getvalue().encode('utf8') == GOLD[f'csv_round{int(round_midi)}'].tobytes()


def test_load_wav_int16_and_resample(tmp_path):
    from scipy.io import wavfile
    t = np.arange(22050) / 22050.0
    x = (0.5 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    wavfile.write(tmp_path / 'a.wav', 44100, (x[:8000] * 32767).astype(np.int16))
    y = batch.load_wav(tmp_path / 'a.wav', 44100)
    assert y.dtype == np.float32 and y.shape == (8000,) and np.abs(y - x[:8000]).max() < 1e-4
    wavfile.write(tmp_path / 'b.wav', 22050, np.stack([x, x], axis=1))
    z = batch.load_wav(tmp_path / 'b.wav', 44100)
    assert z.shape == (44100,) and abs(float(np.abs(z[2000:-2000]).max()) - 0.5) < 0.01
