"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs only in the build container, where
``/root/reference`` is mounted; the GPU box never sees the reference, only these fixtures.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

Shims: a spec'd ``pretty_midi`` stub (extractor.py:23 imports it at module top),
``torchaudio.load`` monkey-patched to return the synthetic wave, and the model weights loaded from
``oracle.model.init_state_dict`` through the reference's own ``load_state_dict`` (extractor.py:108-109).
"""
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True
_pm = types.ModuleType("pretty_midi")
_pm.__spec__ = importlib.machinery.ModuleSpec("pretty_midi", None)
sys.modules["pretty_midi"] = _pm

import torchaudio  # noqa: E402
from etude.config import load_config  # noqa: E402
from etude.data.extractor import AMTAPC_Extractor  # noqa: E402

from etude_b200 import synth  # noqa: E402
from oracle import model as omodel  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def make_extractor(seed):
    sd = omodel.init_state_dict(seed)
    path = "/tmp/etude_golden_sd.pth"
    torch.save(sd, path)
    ex = AMTAPC_Extractor(load_config().extractor, path, device="cpu")
    missing = set(ex.model.state_dict().keys()) - set(sd.keys())
    assert not missing, missing
    return ex, sd


def ref_feature(ex, wave):
    torchaudio.load = lambda p: (torch.from_numpy(np.asarray(wave, np.float32))[None], 16000)
    return ex._wav2feature("synthetic.wav").numpy()


def gen_logmel(ex):
    out = {}
    cases = {
        "noise_1s": synth.noise(16000, 1234),
        "tones_2s": synth.tones(32000, 4321),
        "noise_ragged": synth.noise(16000 + 137, 7),       # N not a multiple of hop
        "noise_short": synth.noise(1025, 9),               # shortest length reflect padding admits (> n_fft/2)
        "silence": np.zeros(4096, np.float32),             # exercises log(0 + 1e-8)
    }
    for k, w in cases.items():
        out[k + "_wave"] = w.astype(np.float32)
        out[k + "_feat"] = ref_feature(ex, w).astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "logmel.npz"), **out)
    print("logmel:", {k: v.shape for k, v in out.items()})


def gen_model(ex):
    g = torch.Generator().manual_seed(99)
    # one window of log-mel-like input: real feature of tones + the -18 padding the driver adds
    feat = ref_feature(ex, synth.tones(256 * 300, 11))                       # 301 frames
    x = torch.from_numpy(omodel.pad_feature(feat)[: 576].T.copy())[None]     # [1, 256, 576]
    with torch.no_grad():
        enc = ex.model.encode(x)
        o = ex.model(x)
    frames = [0, 1, 255, 300, 511]
    out = {
        "input_spec": x.numpy(),
        "enc_sample": enc[0, frames].numpy(),                                # [5, 256, 256]
        "frames": np.array(frames),
        "onset_f": o[0].numpy(), "offset_f": o[1].numpy(), "mpe_f": o[2].numpy(),
        "velocity_f_argmax": o[3].argmax(3).numpy().astype(np.int8),
        "velocity_f_sample": o[3][0, frames].numpy(),
        "attention_sample": o[4][0, frames].numpy(),                         # [5, 4, 88, 256]
        "onset_t": o[5].numpy(), "offset_t": o[6].numpy(), "mpe_t": o[7].numpy(),
        "velocity_t_argmax": o[8].argmax(3).numpy().astype(np.int8),
        "velocity_t_sample": o[8][0, frames].numpy(),
    }
    np.savez_compressed(os.path.join(GOLD, "model_window.npz"), **out)
    print("model:", {k: v.shape for k, v in out.items()})
    del g


def gen_transcript(ex):
    feat = ref_feature(ex, synth.tones(256 * 600 + 19, 21))                  # 601 frames -> 2 windows, ragged tail
    outs = ex._transcript(torch.from_numpy(feat))
    names = ["onset_A", "offset_A", "mpe_A", "velocity_A", "onset_B", "offset_B", "mpe_B", "velocity_B"]
    d = {n: a for n, a in zip(names, outs)}
    d["feature"] = feat
    notes = ex._mpe2note(*outs[4:8], thred_onset=0.5, thred_offset=1.0, thred_mpe=0.5)
    d.update(pack_notes("notes", notes))
    np.savez_compressed(os.path.join(GOLD, "transcript.npz"), **d)
    print("transcript:", {k: v.shape for k, v in d.items()}, "notes", len(notes))


def gen_clip30(ex):
    """BASELINE config 1 (and the hand-off artefact of config 5): the reference's own `extract()` on the 30 s noise clip
    -> extract.json (notes after the min_duration filter, sorted by onset), plus the four B rolls `extract()` used (fp16:
    the parity tolerance on rolls is 2e-2)."""
    import json
    import tempfile
    wave = synth.noise(480000, 1234)
    torchaudio.load = lambda p: (torch.from_numpy(np.asarray(wave, np.float32))[None], 16000)
    feat = ex._wav2feature("synthetic.wav")
    outs = ex._transcript(feat)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "extract.json")
        ex.extract("synthetic.wav", path)
        notes = json.load(open(path))
    d = {"onset_B": outs[4].astype(np.float16), "offset_B": outs[5].astype(np.float16), "mpe_B": outs[6].astype(np.float16),
         "velocity_B": outs[7].astype(np.int8), "n_frames": np.array([feat.shape[0]])}
    d.update(pack_notes("json", notes))
    np.savez_compressed(os.path.join(GOLD, "clip30.npz"), **d)
    print("clip30:", {k: v.shape for k, v in d.items()}, "json notes", len(notes))


def gen_handoff(ex):
    """BASELINE config 5 (infer.py end to end with the reference Structuralize + decoder behind the new extractor): stage 2
    needs madmom / librosa (absent), so the reference's own fixture docs/songs/JPOP16/tempo.json stands for it.  The
    reference chain infer.py:165-210 is run on the reference's extract.json of the 30 s clip (tests/golden/clip30.npz):
    TinyREMITokenizer.encode -> Vocab -> split into bars -> EtudeDecoder.generate (seeded random init, greedy) ->
    decode_to_notes.  Stored: the condition event tokens (what the new extractor must reproduce) and, as evidence that the
    downstream stages accept them, the size of the generated sequence / decoded note list."""
    import json
    import shutil
    import tempfile

    from etude.data.tokenizer import TinyREMITokenizer
    from etude.data.vocab import Vocab
    from etude.models.etude_decoder import EtudeDecoder, EtudeDecoderConfig
    tempo_src = "/root/reference/docs/songs/JPOP16/tempo.json"
    tempo_dst = os.path.join(GOLD, "tempo_JPOP16.json")
    shutil.copyfile(tempo_src, tempo_dst)          # data fixture of the reference (2.4 KB), not source code
    os.chmod(tempo_dst, 0o644)
    z = np.load(os.path.join(GOLD, "clip30.npz"))
    notes = [{"onset": float(a), "offset": float(b), "pitch": int(p), "velocity": int(v)}
             for p, a, b, v in zip(z["json_pitch"], z["json_onset"], z["json_offset"], z["json_velocity"])]
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "extract.json")
        json.dump(notes, open(path, "w"))
        tok = TinyREMITokenizer(tempo_path=tempo_dst)
        events = tok.encode(path)
    tokens = [str(e) for e in events]
    vocab = Vocab()
    vocab.build_from_events([events])
    ids = vocab.encode_sequence(events)
    bars = tok.split_sequence_into_bars(ids, vocab.get_bar_bos_id(), vocab.get_bar_eos_id())
    # the decoder: default architecture, seeded random init, greedy; the first bars with notes are enough to show the chain runs
    torch.manual_seed(0)
    cfg = EtudeDecoderConfig(vocab_size=len(vocab), max_position_embeddings=16384)
    dec = EtudeDecoder(cfg).eval()
    attrs = {"polyphony_bin": 1, "rhythm_intensity_bin": 1, "sustain_bin": 1, "pitch_overlap_bin": 1}
    short = [b[:48] + [b[-1]] if len(b) > 49 else b for b in bars[:3]]
    with torch.no_grad():
        gen = dec.generate(vocab=vocab, all_x_bars=short, target_attributes_per_bar=[attrs] * len(short), temperature=0.0, top_p=0.9,
                           max_output_tokens=96, max_bar_token_limit=32)
    final = tok.decode_to_notes(events=gen, volume_map_path=None) if gen else []
    d = {"tokens": np.array(tokens), "n_bars": np.array([len(bars)]), "n_generated_events": np.array([len(gen)]),
         "n_decoded_notes": np.array([len(final)]), "generated_tokens": np.array([str(e) for e in gen])}
    np.savez_compressed(os.path.join(GOLD, "handoff.npz"), **d)
    print("handoff:", len(tokens), "condition tokens,", len(bars), "bars; decoder generated", len(gen), "events ->", len(final), "notes")


def hft_state_dict(seed=0):
    """The seeded extractor weights with the 128-frame time embedding of HFTConfig (only that tensor depends on num_frame)."""
    sd = dict(omodel.init_state_dict(seed))
    sd["decoder.pos_embedding_time.weight"] = sd["decoder.pos_embedding_time.weight"][:128].clone()
    return sd


def gen_hft(ex):
    """the reference's HFT_Transformer (hft_transformer.py:36-674) on a 4 s clip -- its own pickled-module
    loader, `_wav2feature` (pad_mode="constant"), `_transcript_stride` (n_stride = 32: overlapped 128-frame windows, centre
    halves kept), `_transcript`, and the notes / JSON of `transcribe`."""
    import json
    import pickle
    import tempfile

    from etude.models import amt_apc
    from etude.models.hft_transformer import HFT_Transformer
    cfg = load_config().hft
    m = load_config().extractor.model
    enc = amt_apc.Encoder_SPEC2MIDI(cfg.input.margin_b, cfg.input.num_frame, cfg.feature.n_bins, m.cnn_channel, m.cnn_kernel, m.transformer_hid_dim,
                                    m.encoder_n_layer, m.encoder_n_head, m.transformer_pf_dim, m.dropout, "cpu")
    dec = amt_apc.Decoder_SPEC2MIDI(cfg.input.num_frame, cfg.feature.n_bins, cfg.midi.num_note, cfg.midi.num_velocity, m.transformer_hid_dim,
                                    m.decoder_n_layer, m.decoder_n_head, m.transformer_pf_dim, m.dropout, "cpu")
    model = amt_apc.Model_SPEC2MIDI(enc, dec)
    sd = hft_state_dict(0)
    full = {("encoder_spec2midi." + k[8:] if k.startswith("encoder.") else "decoder_spec2midi." + k[8:]): v for k, v in sd.items()}
    missing, unexpected = model.load_state_dict(full, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    wave = synth.tones(64000, 31)
    torchaudio.load = lambda p: (torch.from_numpy(np.asarray(wave, np.float32))[None], 16000)
    with tempfile.TemporaryDirectory() as td:
        pkl = os.path.join(td, "hft.pkl")
        with open(pkl, "wb") as f:
            pickle.dump(model, f)
        hft = HFT_Transformer(cfg, pkl, device="cpu")
        feat = hft._wav2feature("synthetic.wav")
        stride = hft._transcript_stride(feat, cfg.infer.n_stride, mode="combination")
        plain = hft._transcript(feat, mode="combination")
        out_json = os.path.join(td, "out.json")
        hft.transcribe("synthetic.wav", out_json)
        notes = json.load(open(out_json))
    names = ["onset_A", "offset_A", "mpe_A", "velocity_A", "onset_B", "offset_B", "mpe_B", "velocity_B"]
    d = {"feature": feat.numpy().astype(np.float32)}
    for n, a, b in zip(names, stride, plain):
        d["stride_" + n] = a
        d["plain_" + n] = b
    d.update(pack_notes("json", notes))
    np.savez_compressed(os.path.join(GOLD, "hft.npz"), **d)
    print("hft:", {k: v.shape for k, v in d.items()}, "notes", len(notes))


def pack_notes(prefix, notes):
    return {
        prefix + "_pitch": np.array([n["pitch"] for n in notes], np.int32),
        prefix + "_velocity": np.array([n["velocity"] for n in notes], np.int32),
        prefix + "_onset": np.array([n["onset"] for n in notes], np.float64),
        prefix + "_offset": np.array([n["offset"] for n in notes], np.float64),
    }


def smooth_roll(rng, t, n, width, lo=0.0, hi=1.0):
    x = rng.normal(size=(t + 4 * width, n))
    k = np.hanning(2 * width + 1)
    k /= k.sum()
    y = np.stack([np.convolve(x[:, j], k, mode="same") for j in range(n)], 1)[2 * width : 2 * width + t]
    y = (y - y.min()) / (y.max() - y.min() + 1e-12)
    return (lo + (hi - lo) * y).astype(np.float32)


def notes_cases():
    """Hand-built edge cases + random rolls for _mpe2note."""
    rng = np.random.default_rng(2024)
    cases = {}
    n = 88

    def vel(t):
        return rng.integers(0, 128, size=(t, n)).astype(np.int8)

    t = 400
    cases["smooth"] = (smooth_roll(rng, t, n, 6), smooth_roll(rng, t, n, 6), smooth_roll(rng, t, n, 10), vel(t), 0.5, 0.5, 0.5)
    on = smooth_roll(rng, t, n, 5)
    off = np.minimum(1.0, smooth_roll(rng, t, n, 5) * 1.6).astype(np.float32)      # saturates to exactly 1.0
    cases["offset_saturated"] = (on, off, smooth_roll(rng, t, n, 12), vel(t), 0.5, 1.0, 0.5)
    q = np.round(smooth_roll(rng, t, n, 4) * 8) / 8                                 # plateaus of equal values
    cases["plateaus"] = (q.astype(np.float32), (np.round(smooth_roll(rng, t, n, 4) * 6) / 6).astype(np.float32),
                         (np.round(smooth_roll(rng, t, n, 8) * 4) / 4).astype(np.float32), vel(t), 0.5, 0.5, 0.5)
    cases["white"] = (rng.random((t, n), dtype=np.float32), rng.random((t, n), dtype=np.float32),
                      rng.random((t, n), dtype=np.float32), vel(t), 0.5, 0.5, 0.5)
    z = np.zeros((64, n), np.float32)
    cases["all_zero"] = (z, z, z, vel(64), 0.5, 0.5, 0.5)
    o = np.ones((64, n), np.float32)
    cases["all_one"] = (o, o, o, vel(64), 0.5, 1.0, 0.5)
    cases["one_frame"] = (rng.random((1, n), dtype=np.float32), rng.random((1, n), dtype=np.float32),
                          rng.random((1, n), dtype=np.float32), vel(1), 0.5, 0.5, 0.5)
    cases["two_frames"] = (rng.random((2, n), dtype=np.float32), rng.random((2, n), dtype=np.float32),
                           rng.random((2, n), dtype=np.float32), vel(2), 0.5, 0.5, 0.5)
    v0 = vel(t)
    v0[rng.random((t, n)) < 0.5] = 0                                                # velocity-0 drops
    cases["velocity_zero"] = (smooth_roll(rng, t, n, 3), smooth_roll(rng, t, n, 3), smooth_roll(rng, t, n, 5), v0, 0.4, 0.6, 0.3)
    # sigmoid-like dense rolls as random-init weights produce them (84 % of onset cells >= 0.5)
    sig = lambda a: (1.0 / (1.0 + np.exp(-a))).astype(np.float32)
    cases["dense_sigmoid"] = (sig(rng.normal(1.0, 1.0, (t, n))), sig(rng.normal(12.0, 6.0, (t, n))),
                              sig(rng.normal(0.5, 1.5, (t, n))), vel(t), 0.5, 1.0, 0.5)
    return cases


def gen_notes(ex):
    d = {}
    for name, (on, off, mpe, vel, t_on, t_off, t_mpe) in notes_cases().items():
        for mode_offset in ("shorter", "longer", "offset"):
            if mode_offset != "shorter" and name not in ("smooth", "plateaus"):
                continue
            for mode_velocity in ("ignore_zero", "org"):
                if mode_velocity != "ignore_zero" and name not in ("smooth", "velocity_zero"):
                    continue
                notes = ex._mpe2note(on, off, mpe, vel, thred_onset=t_on, thred_offset=t_off, thred_mpe=t_mpe,
                                     mode_velocity=mode_velocity, mode_offset=mode_offset)
                key = f"{name}__{mode_offset}__{mode_velocity}"
                d.update(pack_notes(key, notes))
                print("notes", key, len(notes))
        d[name + "__onset"], d[name + "__offset"], d[name + "__mpe"], d[name + "__velocity"] = on, off, mpe, vel
        d[name + "__thr"] = np.array([t_on, t_off, t_mpe], np.float64)
    np.savez_compressed(os.path.join(GOLD, "notes.npz"), **d)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    ex, _sd = make_extractor(seed=0)
    only = set(sys.argv[1:])   # e.g. `python oracle/gen_golden.py clip30` regenerates one fixture
    for name, fn in (("logmel", gen_logmel), ("notes", gen_notes), ("model", gen_model), ("transcript", gen_transcript),
                     ("clip30", gen_clip30), ("handoff", gen_handoff), ("hft", gen_hft)):
        if not only or name in only:
            fn(ex)
    print("numpy", np.__version__, "torch", torch.__version__, "torchaudio", torchaudio.__version__)
