"""Round-2 diagnostic (GPU box): per-frame bounds / probs of one 30 s two_head clip from the product path and from the fp32
oracle, dumped for offline analysis of the boundary-cumsum drift (gpurun_out/c2/parity_dump.npz)."""
import sys, tempfile, pathlib
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[2]))
from some_b200 import synth, plugin
from oracle import decode as odecode

out = sys.argv[1]
config = synth.named_config('two_head')
with tempfile.TemporaryDirectory() as d:
    ckpt = synth.write_checkpoint(d, config, seed=1234)
    ins = plugin.MIDIExtractionInference(config=config, model_path=ckpt)
waves = [synth.synth_waveform(900, seconds=30.0), synth.synth_waveform(7, seconds=30.0)]
got = ins.model.infer(waves, return_intermediates=True)
sd = synth.fabricate_state_dict(config, seed=1234)
dump = {}
for i, w in enumerate(waves):
    ref = odecode.infer_clip(sd, config, w, return_intermediates=True)
    g = got[i]
    db = g['bounds'].astype(np.float64) - ref['bounds'].astype(np.float64)
    print(f'clip {i}: bounds err mean {db.mean():+.3e} rms {np.sqrt((db**2).mean()):.3e} max {np.abs(db).max():.3e} '
          f'cumsum drift end {db.sum():+.4f} max |drift| {np.abs(np.cumsum(db)).max():.4f}; '
          f'probs max err {np.abs(g["probs"] - ref["probs"]).max():.3e}; notes {len(g["note_midi"])} vs {len(ref["note_midi"])}')
    dump[f'bounds_gpu{i}'] = g['bounds']; dump[f'bounds_ref{i}'] = ref['bounds']
    dump[f'pmax_gpu{i}'] = g['probs'].max(1); dump[f'pmax_ref{i}'] = ref['probs'].max(1)
    dump[f'parg_gpu{i}'] = g['probs'].argmax(1); dump[f'parg_ref{i}'] = ref['probs'].argmax(1)
np.savez_compressed(out, **dump)
