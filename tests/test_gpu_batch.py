"""GPU tests of the dataset-level driver: many recordings as one var-len batch must give exactly the notes
of the per-recording flow, and batch_infer_dataset must write the CSV that the per-file flow (batch_infer.py:49-81,164-219,
restated in some_b200/batch.py and pinned on the CPU side) produces from those notes."""
import csv
import pathlib
import sys

import numpy as np
import pytest

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / 'golden'))

from slicer_cases import make_case  # noqa: E402
from some_b200 import batch, plugin, synth  # noqa: E402
from some_b200 import slicer as psl  # noqa: E402

pytestmark = pytest.mark.gpu


def _plugin(tmp_path, name='two_head'):
    config = synth.named_config(name)
    ckpt = synth.write_checkpoint(tmp_path, config)
    cls = plugin.QuantizedMIDIExtractionInference if name.startswith('quant') else plugin.MIDIExtractionInference
    return cls(config=config, model_path=ckpt, device='cuda:0'), config


@pytest.mark.parametrize('name', ['two_head', 'quant_two_head'])
def test_many_recordings_one_batch_equals_per_recording(name, tmp_path):
    ins, _ = _plugin(tmp_path, name)
    recs = [make_case('lead_trail'), make_case('short3'), make_case('all_silence'), make_case('tight')[:20 * 44100],
            synth.synth_waveform(31, seconds=12.0, silence_gaps=True)]
    got = batch.transcribe_recordings(ins, recs)
    s = psl.Slicer(44100, max_sil_kept=1000)
    assert len(got) == len(recs)
    for w, (offsets, notes) in zip(recs, got):
        chunks = s.slice(w)                                             # per-recording flow, batch_infer.py:50-54
        assert [c['offset'] for c in chunks] == list(offsets)
        ref = ins.infer([c['waveform'] for c in chunks])
        assert len(ref) == len(notes)
        for a, b in zip(ref, notes):
            for k in ('note_midi', 'note_dur', 'note_rest'):
                np.testing.assert_array_equal(a[k], b[k])


def test_batch_infer_dataset_writes_the_per_file_csv(tmp_path):
    from scipy.io import wavfile
    ins, config = _plugin(tmp_path)
    data = tmp_path / 'raw'
    (data / 'wavs').mkdir(parents=True)
    rng = np.random.default_rng(5)
    rows, waves = [], {}
    for i in range(4):
        w = synth.synth_waveform(40 + i, seconds=float(rng.uniform(6, 14)), silence_gaps=True)
        pcm = np.round(w * 32767).astype(np.int16)
        wavfile.write(data / 'wavs' / f'it{i}.wav', 44100, pcm)
        waves[f'it{i}'] = batch.load_wav(data / 'wavs' / f'it{i}.wav', 44100)
        n_words = int(rng.integers(3, 9))
        ph_num = rng.integers(1, 4, size=n_words)
        durs = np.diff(np.concatenate([[0.0], np.sort(rng.uniform(0, len(w) / 44100, int(ph_num.sum()) - 1)), [len(w) / 44100]]))
        rows.append({'name': f'it{i}', 'ph_seq': ' '.join(['a'] * int(ph_num.sum())), 'ph_dur': ' '.join(f'{d:.6f}' for d in durs),
                     'ph_num': ' '.join(str(int(x)) for x in ph_num)})
    rows.insert(1, {'name': 'absent', 'ph_seq': 'a', 'ph_dur': '0.5', 'ph_num': '1'})
    with open(data / 'transcriptions.csv', 'w', encoding='utf8', newline='') as f:
        wr = csv.DictWriter(f, fieldnames=['name', 'ph_seq', 'ph_dur', 'ph_num'])
        wr.writeheader()
        wr.writerows(rows)
    with pytest.raises(FileExistsError):
        batch.batch_infer_dataset(data, ins, config)
    out = batch.batch_infer_dataset(data, ins, config, csv=tmp_path / 'out.csv', max_frames_per_batch=1500)   # several groups
    got = list(csv.DictReader(open(out, encoding='utf8', newline='')))
    s = psl.Slicer(44100, max_sil_kept=1000)
    assert [r['name'] for r in got] == [r['name'] for r in rows]
    for r in got:
        if r['name'] == 'absent':
            assert r['note_seq'] == '' and r['note_dur'] == ''
            continue
        chunks = s.slice(waves[r['name']])
        notes = ins.infer([c['waveform'] for c in chunks])
        seq, dur = batch.row_notes(r['ph_dur'], r['ph_num'], batch.note_timeline([c['offset'] for c in chunks], notes), False)
        assert (r['note_seq'], r['note_dur']) == (seq, dur)
        assert abs(sum(float(x) for x in dur.split(' ')) - sum(float(x) for x in r['ph_dur'].split(' '))) < 1e-3
