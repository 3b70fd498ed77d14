"""Standard MIDI File output for decoded notes: the step right after the path.

Replaces ``utils.infer_utils.build_midi_file`` (utils/infer_utils.py:79-100) + ``mido.MidiFile.save`` with a dependency-free
writer, so the ``infer.py`` flow (infer.py:41-45) runs where ``mido`` is not installed.

Tick arithmetic follows the reference line by line in meaning: ticks = seconds * tempo * 8 (480 ticks per beat), chunk
offsets rounded to ticks, per-chunk note ends from the rounded CUMULATIVE duration (so rounding never accumulates), a note
clipped at the next chunk's offset, rests and empty notes skipped, note-on delta counted from the previous note-off.

Wire format = SMF as ``mido`` (unpinned in requirements.txt:13; not installed here, so byte parity with mido is restated from
the SMF specification and mido's documented defaults, not pinned): ``MThd`` (format 1, one track, 480 ticks per beat);
``MTrk`` with variable-length delta times, ``FF 51 03`` set_tempo with round(60e6 / bpm) microseconds per beat, note_on
``90 nn 40`` / note_off ``80 nn 40`` on channel 0 with mido's default velocity 64, and the ``FF 2F 00`` end-of-track that
mido appends on save.  (Running status would never apply: note_on and note_off statuses alternate.)
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np

TICKS_PER_BEAT = 480
DEFAULT_VELOCITY = 64


def note_events(offsets: Sequence[float], segments: Sequence[Dict[str, np.ndarray]], tempo: float = 120) -> List[Tuple[int, int, int]]:
    """[(start_tick, end_tick, midi_note)] of the voiced notes, in track order (infer_utils.py:83-98)."""
    ticks_per_second = tempo * 8
    chunk_tick = [round(o * ticks_per_second) for o in offsets]
    events: List[Tuple[int, int, int]] = []
    for i, (t0, seg) in enumerate(zip(chunk_tick, segments)):
        pitch = np.round(seg['note_midi']).astype(np.int64).tolist()
        ends = np.round(np.cumsum(seg['note_dur']) * ticks_per_second).astype(np.int64)
        length = np.diff(ends, prepend=0).tolist()
        rest = seg['note_rest'].tolist()
        limit = chunk_tick[i + 1] if i < len(chunk_tick) - 1 else None
        start = t0
        for p, n, r in zip(pitch, length, rest):
            end = start + n
            if limit is not None and end > limit:
                end = limit
            if start < end and not r:
                events.append((int(start), int(end), int(p)))
            start = end
    return events


def _varlen(value: int) -> bytes:
    if value < 0:
        raise ValueError('message time must be non-negative in a MIDI file')   # mido raises on negative delta times too
    out = [value & 0x7f]
    value >>= 7
    while value:
        out.append((value & 0x7f) | 0x80)
        value >>= 7
    return bytes(reversed(out))


def bpm2tempo(bpm: float) -> int:
    return int(round(60 * 1000000 / bpm))


class MidiFile:
    """The little of mido.MidiFile that infer.py uses: ``tracks`` (one list of (delta, kind, data) messages) and ``save``."""

    def __init__(self, tempo: float = 120, ticks_per_beat: int = TICKS_PER_BEAT):
        self.type = 1
        self.ticks_per_beat = ticks_per_beat
        self.tracks: List[List[Tuple[int, str, Tuple[int, ...]]]] = [[(0, 'set_tempo', (bpm2tempo(tempo),))]]

    def add_note(self, delta_on: int, length: int, note: int):
        if not 0 <= note <= 127:
            raise ValueError(f'attribute must be in range 0..127: note={note}')
        self.tracks[0].append((int(delta_on), 'note_on', (note, DEFAULT_VELOCITY)))
        self.tracks[0].append((int(length), 'note_off', (note, DEFAULT_VELOCITY)))

    def to_bytes(self) -> bytes:
        chunks = [b'MThd' + struct.pack('>IHHH', 6, self.type, len(self.tracks), self.ticks_per_beat)]
        for track in self.tracks:
            data = bytearray()
            for delta, kind, args in track:
                data += _varlen(delta)
                if kind == 'set_tempo':
                    data += b'\xff\x51\x03' + struct.pack('>I', args[0])[1:]
                elif kind == 'note_on':
                    data += bytes((0x90, args[0], args[1]))
                elif kind == 'note_off':
                    data += bytes((0x80, args[0], args[1]))
                else:
                    raise ValueError(kind)
            data += b'\x00\xff\x2f\x00'                                      # end_of_track, delta 0
            chunks.append(b'MTrk' + struct.pack('>I', len(data)) + bytes(data))
        return b''.join(chunks)

    def save(self, filename):
        with open(os.fspath(filename), 'wb') as f:
            f.write(self.to_bytes())


def build_midi_file(offsets: List[float], segments: List[Dict[str, np.ndarray]], tempo=120) -> MidiFile:
    """Same call as utils.infer_utils.build_midi_file; returns an object with ``save(path)``."""
    midi = MidiFile(tempo=tempo)
    last = 0
    for start, end, note in note_events(offsets, segments, tempo):
        midi.add_note(start - last, end - start, note)
        last = end
    return midi


def parse_midi(data: bytes):
    """Minimal SMF reader for the tests: returns (format, ticks_per_beat, [[(delta, status, data bytes)]])."""
    assert data[:4] == b'MThd' and struct.unpack('>I', data[4:8])[0] == 6
    fmt, ntrk, tpb = struct.unpack('>HHH', data[8:14])
    pos, tracks = 14, []
    for _ in range(ntrk):
        assert data[pos:pos + 4] == b'MTrk'
        ln = struct.unpack('>I', data[pos + 4:pos + 8])[0]
        body, p, msgs = data[pos + 8:pos + 8 + ln], 0, []
        while p < len(body):
            delta = 0
            while True:
                b = body[p]
                p += 1
                delta = (delta << 7) | (b & 0x7f)
                if not b & 0x80:
                    break
            status = body[p]
            if status == 0xff:
                kind, n = body[p + 1], body[p + 2]
                msgs.append((delta, (0xff, kind), bytes(body[p + 3:p + 3 + n])))
                p += 3 + n
            else:
                msgs.append((delta, status, bytes(body[p + 1:p + 3])))
                p += 3
        tracks.append(msgs)
        pos += 8 + ln
    return fmt, tpb, tracks
