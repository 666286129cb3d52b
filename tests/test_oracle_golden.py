"""The oracle pinned against outputs of the reference itself (tests/golden, made by oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import NOTE_CASES, note_variants, unpack_notes
from oracle import logmel as ologmel
from oracle import model as omodel
from oracle import notes as onotes


@pytest.mark.parametrize("case", ["noise_1s", "tones_2s", "noise_ragged", "noise_short", "silence"])
def test_logmel_oracle_matches_reference(golden, case):
    z = golden("logmel")
    feat = ologmel.logmel(z[case + "_wave"])
    ref = z[case + "_feat"]
    assert feat.shape == ref.shape == (1 + len(z[case + "_wave"]) // 256, 256)
    # float64 restatement vs torchaudio fp32: tolerance 1e-3 in the log domain; silence is exact log(1e-8)
    assert np.abs(feat - ref).max() <= 1e-3


def test_mel_filterbank_structure():
    fb = ologmel.mel_filterbank()
    assert fb.shape == (1025, 256)
    assert (fb >= 0).all() and (fb.max(0) > 0).all()
    assert int((fb > 1e-12).sum()) == 2036 and int((fb > 1e-12).sum(1).max()) <= 2   # SURVEY K3 probe


@pytest.mark.parametrize("case", NOTE_CASES)
def test_notes_oracle_bit_exact(golden, case):
    z = golden("notes")
    t_on, t_off, t_mpe = z[case + "__thr"]
    for mo, mv, ref in note_variants(z, case):
        got = onotes.mpe2note(z[case + "__onset"], z[case + "__offset"], z[case + "__mpe"], z[case + "__velocity"],
                              thred_onset=t_on, thred_offset=t_off, thred_mpe=t_mpe, mode_velocity=mv, mode_offset=mo)
        assert got == ref, (case, mo, mv)


def test_model_oracle_matches_reference(golden):
    z = golden("model_window")
    sd = omodel.init_state_dict(0)
    x = torch.from_numpy(z["input_spec"])
    with torch.no_grad():
        enc = omodel.encode(sd, x)
        o = omodel.decode(sd, enc)
    fr = z["frames"]
    # fp32 on both sides; the tolerances only absorb the summation order of the host's BLAS kernels (the golden
    # vectors were made on another CPU: AVX-512 vs AVX2 code paths differ by ~1e-4 on the O(10) encoder activations)
    assert np.abs(enc[0, fr].numpy() - z["enc_sample"]).max() <= 1e-3
    for i, k in [(0, "onset_f"), (1, "offset_f"), (2, "mpe_f"), (5, "onset_t"), (6, "offset_t"), (7, "mpe_t")]:
        assert np.abs(o[i].numpy() - z[k]).max() <= 2e-5, k
    assert np.abs(o[3][0, fr].numpy() - z["velocity_f_sample"]).max() <= 1e-4
    assert np.abs(o[8][0, fr].numpy() - z["velocity_t_sample"]).max() <= 1e-4
    assert np.abs(o[4][0, fr].numpy() - z["attention_sample"]).max() <= 1e-5
    assert (o[3].argmax(3).numpy() == z["velocity_f_argmax"]).mean() >= 0.999
    assert (o[8].argmax(3).numpy() == z["velocity_t_argmax"]).mean() >= 0.999


def test_transcript_oracle_matches_reference(golden):
    z = golden("transcript")
    sd = omodel.init_state_dict(0)
    outs = omodel.transcript(sd, z["feature"], batch=2)
    names = ["onset_A", "offset_A", "mpe_A", "velocity_A", "onset_B", "offset_B", "mpe_B", "velocity_B"]
    for n, a in zip(names, outs):
        assert a.shape == z[n].shape == (1024, 88) and a.dtype == z[n].dtype
        if a.dtype == np.int8:
            assert (a == z[n]).mean() >= 0.999, n
        else:
            assert np.abs(a - z[n]).max() <= 2e-5, n
    # notes: bit-exact when the note stage is fed the reference's own rolls
    got = onotes.mpe2note(z["onset_B"], z["offset_B"], z["mpe_B"], z["velocity_B"], thred_onset=0.5, thred_offset=1.0,
                          thred_mpe=0.5)
    assert got == unpack_notes(z, "notes")


def test_oracle_config1_clip30_matches_reference_extract(golden):
    """BASELINE config 1: the oracle's whole path (log-mel -> 4 windows -> notes -> min_duration filter -> onset order) on the
    30 s noise clip against the reference's own `extract()` output (tests/golden/clip30.npz)."""
    from etude_b200 import synth
    from oracle import logmel as ologmel
    z = golden("clip30")
    sd = omodel.init_state_dict(0)
    feat = ologmel.logmel(synth.noise(480000, 1234), dtype=np.float32)
    assert feat.shape == (int(z["n_frames"][0]), 256)
    outs = omodel.transcript(sd, feat, batch=4)
    for name, got in zip(("onset_B", "offset_B", "mpe_B"), outs[4:7]):
        assert np.abs(got - z[name].astype(np.float32)).max() <= 1.5e-3, name   # fp16 fixture + fp64-vs-fp32 log-mel
    assert (outs[7] == z["velocity_B"]).mean() >= 0.995
    notes = onotes.mpe2note(outs[4], outs[5], outs[6], outs[7], thred_onset=0.5, thred_offset=1.0, thred_mpe=0.5)
    kept = sorted((n for n in notes if not (n["offset"] - n["onset"] < 0.08)), key=lambda n: n["onset"])
    ref = unpack_notes(z, "json")
    key = lambda n: (n["pitch"], round(n["onset"] / 0.016))
    a, b = {key(n) for n in kept}, {key(n) for n in ref}
    f1 = 2 * len(a & b) / max(1, len(a) + len(b))
    assert f1 >= 0.97, (f1, len(kept), len(ref))


@pytest.mark.parametrize("rates", [(44100, 16000), (48000, 16000), (22050, 16000), (8000, 16000), (16000, 16000)])
def test_resample_oracle_matches_torchaudio(rates):
    """The ingest oracle (channel mean + sinc resampling, reference extractor.py:181-184) against the installed torchaudio:
    polyphase kernel bit-identical, output within fp32 accumulation noise, length = ceil(new * n / orig)."""
    import math

    import torch
    import torchaudio
    import torchaudio.functional.functional as TF
    from oracle import resample as oresample
    orig, new = rates
    rng = np.random.default_rng(orig)
    x = rng.uniform(-0.5, 0.5, (2, orig // 3 + 77)).astype(np.float32)
    ref = torchaudio.transforms.Resample(orig, new)(torch.mean(torch.from_numpy(x), dim=0)).numpy()
    got = oresample.resample(x, orig, new)
    assert got.shape == ref.shape == (math.ceil(new * x.shape[1] / orig),)
    assert np.abs(got - ref).max() <= 1e-6
    if orig != new:
        k, w = TF._get_sinc_resample_kernel(orig, new, math.gcd(orig, new))
        mine, w2, _, _ = oresample.sinc_kernel(orig, new)
        assert w == w2 and np.array_equal(k[:, 0].numpy(), mine)
