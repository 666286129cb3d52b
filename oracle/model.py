"""Oracle: hFT-Transformer extractor model, torch fp32 on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  A functional restatement, from a
plain ``state_dict``, of the reference modules

    Model_SPEC2MIDI        etude/models/amt_apc.py:23-49
    Encoder_SPEC2MIDI      etude/models/amt_apc.py:55-120
    Decoder_SPEC2MIDI      etude/models/amt_apc.py:126-230
    EncoderLayer           etude/models/amt_apc.py:236-259
    DecoderLayer_Zero      etude/models/amt_apc.py:261-286
    DecoderLayer           etude/models/amt_apc.py:288-320
    MultiHeadAttention     etude/models/amt_apc.py:322-374
    PositionwiseFFN        etude/models/amt_apc.py:376-392

and of the sliding-window driver ``AMTAPC_Extractor._transcript``
(etude/data/extractor.py:199-253).  Dropout is identity (eval mode).  It is a
floating-point path, so the oracle is a torch fp32 reference, as the task
allows; it is pinned against the real reference modules by
``oracle/gen_golden.py`` (max-abs 0 expected: same ATen ops, same order).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# default ExtractorConfig constants (etude/config/schema.py:68-121)
N_BIN = 256
N_FRAME = 512
MARGIN = 32
N_PROC = 2 * MARGIN + 1
N_NOTE = 88
N_VELOCITY = 128
HID = 256
N_HEADS = 4
HEAD_DIM = HID // N_HEADS
MIN_VALUE = -18.0


def init_state_dict(seed=0):
    """Seeded random-init state_dict with the reference's key names/shapes.

    Mirrors what ``_load_model`` yields for an empty checkpoint
    (etude/data/extractor.py:78-113, strict=False): default nn.Linear /
    nn.Conv2d / nn.Embedding / nn.LayerNorm initialisation.  Built from plain
    torch.nn layers (no reference import) so it can run on the GPU box; the
    construction order differs from the reference's, so the VALUES differ
    from ``torch.manual_seed(seed)`` + reference constructor -- goldens carry
    their own state_dict.
    """
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
        return w, b

    sd = {}

    def put_lin(prefix, out_f, in_f):
        w, b = lin(out_f, in_f)
        sd[prefix + ".weight"] = w
        sd[prefix + ".bias"] = b

    def put_mha(prefix):
        for n in ("fc_q", "fc_k", "fc_v", "fc_o"):
            put_lin(f"{prefix}.{n}", HID, HID)

    def put_ln(prefix):
        # perturbed away from (1, 0) so LN affine terms are exercised
        sd[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(HID, generator=g)
        sd[prefix + ".bias"] = 0.1 * torch.randn(HID, generator=g)

    def put_ffn(prefix):
        put_lin(prefix + ".fc_1", 512, HID)
        put_lin(prefix + ".fc_2", HID, 512)

    bound = 1.0 / math.sqrt(5.0)
    sd["encoder.conv.weight"] = (torch.rand(4, 1, 1, 5, generator=g) * 2 - 1) * bound
    sd["encoder.conv.bias"] = (torch.rand(4, generator=g) * 2 - 1) * bound
    put_lin("encoder.tok_embedding_freq", HID, 244)
    sd["encoder.pos_embedding_freq.weight"] = torch.randn(N_BIN, HID, generator=g)
    for i in range(3):
        p = f"encoder.layers_freq.{i}"
        put_ln(p + ".layer_norm")
        put_mha(p + ".self_attention")
        put_ffn(p + ".positionwise_feedforward")
    sd["decoder.pos_embedding_freq.weight"] = torch.randn(N_NOTE, HID, generator=g)
    p = "decoder.layer_zero_freq"
    put_ln(p + ".layer_norm")
    put_mha(p + ".encoder_attention")
    put_ffn(p + ".positionwise_feedforward")
    for i in range(2):
        p = f"decoder.layers_freq.{i}"
        put_ln(p + ".layer_norm")
        put_mha(p + ".self_attention")
        put_mha(p + ".encoder_attention")
        put_ffn(p + ".positionwise_feedforward")
    for dom in ("freq", "time"):
        for n in ("onset", "offset", "mpe"):
            put_lin(f"decoder.fc_{n}_{dom}", 1, HID)
        put_lin(f"decoder.fc_velocity_{dom}", N_VELOCITY, HID)
    sd["decoder.pos_embedding_time.weight"] = torch.randn(N_FRAME, HID, generator=g)
    for i in range(3):
        p = f"decoder.layers_time.{i}"
        put_ln(p + ".layer_norm")
        put_mha(p + ".self_attention")
        put_ffn(p + ".positionwise_feedforward")
    return sd


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def mha(sd, p, query, key, value):
    """MultiHeadAttentionLayer.forward (amt_apc.py:336-374)."""
    b = query.shape[0]
    q = _lin(sd, p + ".fc_q", query).view(b, -1, N_HEADS, HEAD_DIM).permute(0, 2, 1, 3)
    k = _lin(sd, p + ".fc_k", key).view(b, -1, N_HEADS, HEAD_DIM).permute(0, 2, 1, 3)
    v = _lin(sd, p + ".fc_v", value).view(b, -1, N_HEADS, HEAD_DIM).permute(0, 2, 1, 3)
    energy = torch.matmul(q, k.permute(0, 1, 3, 2)) / math.sqrt(HEAD_DIM)
    attention = torch.softmax(energy, dim=-1)
    x = torch.matmul(attention, v).permute(0, 2, 1, 3).contiguous().view(b, -1, HID)
    return _lin(sd, p + ".fc_o", x), attention


def _ln(sd, p, x):
    return F.layer_norm(x, (HID,), sd[p + ".layer_norm.weight"], sd[p + ".layer_norm.bias"], 1e-5)


def _ffn(sd, p, x):
    """PositionwiseFeedforwardLayer.forward (amt_apc.py:383-392)."""
    return _lin(sd, p + ".positionwise_feedforward.fc_2",
                torch.relu(_lin(sd, p + ".positionwise_feedforward.fc_1", x)))


def encoder_layer(sd, p, src):
    """EncoderLayer.forward (amt_apc.py:244-259): one LN module used twice."""
    a, _ = mha(sd, p + ".self_attention", src, src, src)
    src = _ln(sd, p, src + a)
    return _ln(sd, p, src + _ffn(sd, p, src))


def decoder_layer_zero(sd, p, enc, trg):
    """DecoderLayer_Zero.forward (amt_apc.py:269-286)."""
    a, att = mha(sd, p + ".encoder_attention", trg, enc, enc)
    trg = _ln(sd, p, trg + a)
    return _ln(sd, p, trg + _ffn(sd, p, trg)), att


def decoder_layer(sd, p, enc, trg):
    """DecoderLayer.forward (amt_apc.py:297-320)."""
    a, _ = mha(sd, p + ".self_attention", trg, trg, trg)
    trg = _ln(sd, p, trg + a)
    a, att = mha(sd, p + ".encoder_attention", trg, enc, enc)
    trg = _ln(sd, p, trg + a)
    return _ln(sd, p, trg + _ffn(sd, p, trg)), att


def embed(sd, spec_in):
    """Encoder front part (amt_apc.py:74-111): unfold -> conv(1x5) -> Linear(244,256) -> *16 + pos."""
    b = spec_in.shape[0]
    spec = spec_in.unfold(2, N_PROC, 1).permute(0, 2, 1, 3).contiguous()
    x = spec.reshape(b * N_FRAME, N_BIN, N_PROC).unsqueeze(1)
    x = F.conv2d(x, sd["encoder.conv.weight"], sd["encoder.conv.bias"]).permute(0, 2, 1, 3).contiguous()
    x = x.reshape(b * N_FRAME, N_BIN, -1)
    x = _lin(sd, "encoder.tok_embedding_freq", x)
    return x * math.sqrt(HID) + sd["encoder.pos_embedding_freq.weight"][None]


def encode(sd, spec_in):
    """Encoder_SPEC2MIDI.forward (amt_apc.py:74-120) -> [B, 512, 256, 256]."""
    b = spec_in.shape[0]
    x = embed(sd, spec_in)
    for i in range(3):
        x = encoder_layer(sd, f"encoder.layers_freq.{i}", x)
    return x.reshape(b, N_FRAME, N_BIN, HID)


def _heads(sd, dom, x):
    return (torch.sigmoid(_lin(sd, f"decoder.fc_onset_{dom}", x)),
            torch.sigmoid(_lin(sd, f"decoder.fc_offset_{dom}", x)),
            torch.sigmoid(_lin(sd, f"decoder.fc_mpe_{dom}", x)),
            _lin(sd, f"decoder.fc_velocity_{dom}", x))


def decode(sd, enc, return_intermediates=False):
    """Decoder_SPEC2MIDI.forward (amt_apc.py:159-230) -> the 9-tuple."""
    b = enc.shape[0]
    enc = enc.reshape(b * N_FRAME, N_BIN, HID)
    trg = sd["decoder.pos_embedding_freq.weight"][None].repeat(b * N_FRAME, 1, 1)
    trg, att = decoder_layer_zero(sd, "decoder.layer_zero_freq", enc, trg)
    for i in range(2):
        trg, att = decoder_layer(sd, f"decoder.layers_freq.{i}", enc, trg)
    att = att.reshape(b, N_FRAME, N_HEADS, N_NOTE, N_BIN)
    on_f, off_f, mpe_f, vel_f = _heads(sd, "freq", trg)
    on_f = on_f.reshape(b, N_FRAME, N_NOTE)
    off_f = off_f.reshape(b, N_FRAME, N_NOTE)
    mpe_f = mpe_f.reshape(b, N_FRAME, N_NOTE)
    vel_f = vel_f.reshape(b, N_FRAME, N_NOTE, N_VELOCITY)
    midi_freq = trg
    t = trg.reshape(b, N_FRAME, N_NOTE, HID).permute(0, 2, 1, 3).contiguous().reshape(b * N_NOTE, N_FRAME, HID)
    t = t * math.sqrt(HID) + sd["decoder.pos_embedding_time.weight"][None]
    for i in range(3):
        t = encoder_layer(sd, f"decoder.layers_time.{i}", t)
    on_t, off_t, mpe_t, vel_t = _heads(sd, "time", t)
    on_t = on_t.reshape(b, N_NOTE, N_FRAME).permute(0, 2, 1).contiguous()
    off_t = off_t.reshape(b, N_NOTE, N_FRAME).permute(0, 2, 1).contiguous()
    mpe_t = mpe_t.reshape(b, N_NOTE, N_FRAME).permute(0, 2, 1).contiguous()
    vel_t = vel_t.reshape(b, N_NOTE, N_FRAME, N_VELOCITY).permute(0, 2, 1, 3).contiguous()
    out = (on_f, off_f, mpe_f, vel_f, att, on_t, off_t, mpe_t, vel_t)
    if return_intermediates:
        return out, {"midi_freq": midi_freq, "midi_time": t}
    return out


@torch.no_grad()
def forward(sd, input_spec):
    """Model_SPEC2MIDI.forward (amt_apc.py:29-49): [B,256,576] -> 9-tuple."""
    return decode(sd, encode(sd, input_spec))


def pad_feature(feature):
    """The -18 padding of ``_transcript`` (extractor.py:210-213) -> [32 + T_pad + 32, 256]."""
    a = np.asarray(feature, dtype=np.float32)
    t = a.shape[0]
    len_s = int(np.ceil(t / N_FRAME) * N_FRAME) - t
    return np.concatenate([np.full((MARGIN, N_BIN), MIN_VALUE, np.float32), a,
                           np.full((len_s + MARGIN, N_BIN), MIN_VALUE, np.float32)], axis=0)


@torch.no_grad()
def transcript(sd, feature, batch=1):
    """AMTAPC_Extractor._transcript (extractor.py:199-253), mode="combination".

    Returns the 8 arrays (onset_A, offset_A, mpe_A, velocity_A, onset_B,
    offset_B, mpe_B, velocity_B), each with T_pad rows (tail NOT trimmed).
    ``batch`` > 1 stacks windows (results identical per window; used to make
    the CPU baseline use its cores well).
    """
    a_in = torch.from_numpy(pad_feature(feature))
    t = np.asarray(feature).shape[0]
    t_pad = a_in.shape[0] - 2 * MARGIN
    outs = [np.zeros((t_pad, N_NOTE), np.float32) for _ in range(3)] + [np.zeros((t_pad, N_NOTE), np.int8)]
    outs = outs + [np.zeros((t_pad, N_NOTE), np.float32) for _ in range(3)] + [np.zeros((t_pad, N_NOTE), np.int8)]
    starts = list(range(0, t, N_FRAME))
    for s in range(0, len(starts), batch):
        chunk = starts[s : s + batch]
        spec = torch.stack([a_in[i : i + N_FRAME + 2 * MARGIN].T for i in chunk], 0)
        o = forward(sd, spec)
        for bi, i in enumerate(chunk):
            outs[0][i : i + N_FRAME] = o[0][bi].numpy()
            outs[1][i : i + N_FRAME] = o[1][bi].numpy()
            outs[2][i : i + N_FRAME] = o[2][bi].numpy()
            outs[3][i : i + N_FRAME] = o[3][bi].argmax(2).numpy()
            outs[4][i : i + N_FRAME] = o[5][bi].numpy()
            outs[5][i : i + N_FRAME] = o[6][bi].numpy()
            outs[6][i : i + N_FRAME] = o[7][bi].numpy()
            outs[7][i : i + N_FRAME] = o[8][bi].argmax(2).numpy()
    return tuple(outs)
