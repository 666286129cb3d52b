"""HFT_Transformer: the pickled-checkpoint loader on the CPU; on the GPU the whole class against
tests/golden/hft.npz, which oracle/gen_golden.py made with the reference's own HFT_Transformer (hft_transformer.py) on a 4 s clip."""
import json
import pickle
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import report, unpack_notes


def _seeded_hft_state_dict():
    from oracle import model as omodel
    sd = dict(omodel.init_state_dict(0))
    sd["decoder.pos_embedding_time.weight"] = sd["decoder.pos_embedding_time.weight"][:128].clone()
    return sd


def _pickle_like_the_original(sd, path, module_name="model_spec2midi_test"):
    """A checkpoint shaped like the hFT-Transformer release: a pickled nn.Module tree whose classes live in a `model*` module
    that does not exist when the file is read back."""
    mod = types.ModuleType(module_name)
    sys.modules[module_name] = mod

    def cls(name):
        c = type(name, (nn.Module,), {"__module__": module_name})
        setattr(mod, name, c)
        return c
    Model, Enc, Dec, Leaf = cls("Model_SPEC2MIDI"), cls("Encoder_SPEC2MIDI"), cls("Decoder_SPEC2MIDI"), cls("Leaf")
    model, enc, dec = Model(), Enc(), Dec()
    model.encoder_spec2midi, model.decoder_spec2midi = enc, dec
    for key, val in sd.items():
        root = enc if key.startswith("encoder.") else dec
        parts = key.split(".")[1:]
        m = root
        for p in parts[:-1]:
            if p not in m._modules:
                m.add_module(p, Leaf())
            m = m._modules[p]
        m.register_parameter(parts[-1], nn.Parameter(val.clone(), requires_grad=False))
    enc.scale_freq = torch.sqrt(torch.FloatTensor([256]))          # plain tensor attributes, as in amt_apc.py:72
    with open(path, "wb") as f:
        pickle.dump(model, f)
    del sys.modules[module_name]


def test_pickled_checkpoint_loads_without_the_original_classes(tmp_path):
    from etude_b200 import hft
    from etude_b200.weights import N_WEIGHT_FLOATS_HFT, pack_state_dict
    sd = _seeded_hft_state_dict()
    path = tmp_path / "hft.pkl"
    _pickle_like_the_original(sd, path)
    with pytest.raises(Exception):
        pickle.load(open(path, "rb"))                                # the plain unpickler cannot resolve `model_spec2midi_test`
    got = hft.load_pickled_state_dict(path)
    assert set(got) == set(sd)
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    blob, missing = pack_state_dict(got, strict=True, n_frame=128)
    assert blob.size == N_WEIGHT_FLOATS_HFT and not missing
    # a 512-frame checkpoint is rejected with a shape error, not silently cropped
    sd512 = dict(sd)
    sd512["decoder.pos_embedding_time.weight"] = torch.zeros(512, 256)
    with pytest.raises(ValueError):
        pack_state_dict(sd512, strict=True, n_frame=128)


def test_hft_config_defaults_and_errors(tmp_path):
    from etude_b200 import hft
    cfg = hft.HFTConfig()
    assert (cfg.input.num_frame, cfg.input.min_value, cfg.feature.pad_mode, cfg.infer.n_stride) == (128, -80.0, "constant", 32)
    assert (cfg.infer.thred_onset, cfg.infer.thred_offset, cfg.infer.thred_mpe) == (0.75, 0.5, 0.5)
    bad = hft.HFTConfig()
    bad.input.num_frame = 256
    with pytest.raises(ValueError):
        hft.validate(bad)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            hft.HFT_Transformer(cfg, tmp_path / "none.pkl", device="auto")


@pytest.mark.gpu
def test_hft_transformer_vs_reference(golden, tmp_path, monkeypatch):
    """_wav2feature (constant padding), _transcript_stride (device-side stitching of the overlapped windows), _transcript and
    transcribe against the reference class: log-mel <= 1e-3, rolls <= 2e-2, velocity agreement >= 98 %."""
    import torchaudio

    from etude_b200 import hft, synth
    z = golden("hft")
    path = tmp_path / "hft.pkl"
    _pickle_like_the_original(_seeded_hft_state_dict(), path)
    cfg = hft.HFTConfig()
    tr = hft.HFT_Transformer(cfg, path, device="cuda:0")
    assert tr.engine.n_frame == 128
    wave = synth.tones(64000, 31)
    monkeypatch.setattr(torchaudio, "load", lambda p: (torch.from_numpy(wave)[None], 16000))
    feat = tr._wav2feature("x.wav")
    assert tuple(feat.shape) == z["feature"].shape == (251, 256)
    e_feat = float(np.abs(feat.numpy() - z["feature"]).max())
    assert e_feat <= 1e-3
    names = ["onset_A", "offset_A", "mpe_A", "velocity_A", "onset_B", "offset_B", "mpe_B", "velocity_B"]
    vals = {"logmel_maxabs": e_feat}
    for kind, outs in (("stride", tr._transcript_stride(z["feature"], cfg.infer.n_stride)), ("plain", tr._transcript(z["feature"]))):
        assert len(outs) == 8
        for n, a in zip(names, outs):
            ref = z[f"{kind}_{n}"]
            assert a.shape == ref.shape == (256, 88) and a.dtype == ref.dtype, (kind, n)
            vals[f"{kind}_{n}"] = float((a == ref).mean()) if a.dtype == np.int8 else float(np.abs(a - ref).max())
    report("hft_transformer_vs_reference (maxabs of the rolls, agreement of the int8 velocities)", **vals)
    for k, v in vals.items():
        if "velocity" in k:
            assert v >= 0.98, (k, v)
        elif k != "logmel_maxabs":
            assert v <= 2e-2, (k, v)
    # frequency-axis-only mode and the JSON of transcribe()
    assert len(tr._transcript_stride(z["feature"], 32, mode="single")) == 4
    out = tmp_path / "notes.json"
    tr.transcribe("x.wav", out)
    notes = json.loads(out.read_text())
    ref = unpack_notes(z, "json")
    assert all(list(n.keys()) == ["pitch", "onset", "offset", "velocity"] for n in notes)
    key = lambda n: (n["pitch"], round(n["onset"] / 0.016))
    a, b = {key(n) for n in notes}, {key(n) for n in ref}
    report("hft_transcribe_json", notes_ours=len(notes), notes_reference=len(ref), common=len(a & b))
    assert abs(len(notes) - len(ref)) <= 2
    # notes of the stage itself are bit-exact on identical rolls: the reference's rolls through our _mpe2note
    from oracle import notes as onotes
    args = (z["stride_onset_B"], z["stride_offset_B"], z["stride_mpe_B"], z["stride_velocity_B"])
    assert tr._mpe2note(*args, thred_onset=0.5, thred_offset=0.5, thred_mpe=0.5) == onotes.mpe2note(*args, 0.5, 0.5, 0.5)
