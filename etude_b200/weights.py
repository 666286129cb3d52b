"""Checkpoint layout of the extractor: the reference's state_dict keys (etude/data/extractor.py:78-113,
SURVEY.md A14) and their flattening into the fp32 blob ``etude_create`` consumes.

``checkpoints/extractor/latest.pth`` loads unchanged: keys and shapes are the reference module's; missing keys
keep the seeded default initialisation (the reference loads with ``strict=False``).
"""
import math

import numpy as np
import torch

HID, PF, N_BIN, N_NOTE, N_VEL, N_FRAME = 256, 512, 256, 88, 128, 512


def _linear(prefix, out_f, in_f):
    return [(prefix + ".weight", (out_f, in_f)), (prefix + ".bias", (out_f,))]


def _mha(prefix):
    out = []
    for n in ("fc_q", "fc_k", "fc_v", "fc_o"):
        out += _linear(f"{prefix}.{n}", HID, HID)
    return out


def _ln(prefix):
    return [(prefix + ".layer_norm.weight", (HID,)), (prefix + ".layer_norm.bias", (HID,))]


def _ffn(prefix):
    return _linear(prefix + ".positionwise_feedforward.fc_1", PF, HID) + _linear(prefix + ".positionwise_feedforward.fc_2", HID, PF)


def layout(n_frame=N_FRAME):
    """(key, shape) list for a window length of ``n_frame`` frames: 512 for the AMT-APC extractor (ExtractorConfig), 128 for
    HFT_Transformer (HFTConfig); only ``decoder.pos_embedding_time`` depends on it (amt_apc.py:149)."""
    out = [("encoder.conv.weight", (4, 1, 1, 5)), ("encoder.conv.bias", (4,))]
    out += _linear("encoder.tok_embedding_freq", HID, 244)
    out += [("encoder.pos_embedding_freq.weight", (N_BIN, HID))]
    for i in range(3):
        p = f"encoder.layers_freq.{i}"
        out += _ln(p) + _mha(p + ".self_attention") + _ffn(p)
    out += [("decoder.pos_embedding_freq.weight", (N_NOTE, HID))]
    p = "decoder.layer_zero_freq"
    out += _ln(p) + _mha(p + ".encoder_attention") + _ffn(p)
    for i in range(2):
        p = f"decoder.layers_freq.{i}"
        out += _ln(p) + _mha(p + ".self_attention") + _mha(p + ".encoder_attention") + _ffn(p)
    for n in ("onset", "offset", "mpe"):
        out += _linear(f"decoder.fc_{n}_freq", 1, HID)
    out += _linear("decoder.fc_velocity_freq", N_VEL, HID)
    out += [("decoder.pos_embedding_time.weight", (int(n_frame), HID))]
    for i in range(3):
        p = f"decoder.layers_time.{i}"
        out += _ln(p) + _mha(p + ".self_attention") + _ffn(p)
    for n in ("onset", "offset", "mpe"):
        out += _linear(f"decoder.fc_{n}_time", 1, HID)
    out += _linear("decoder.fc_velocity_time", N_VEL, HID)
    return out


#: (key, shape) in blob order == the reference module's state_dict() order
STATE_DICT_LAYOUT = layout()
N_WEIGHT_FLOATS = sum(int(np.prod(s)) for _, s in STATE_DICT_LAYOUT)
N_WEIGHT_FLOATS_HFT = sum(int(np.prod(s)) for _, s in layout(128))
assert N_WEIGHT_FLOATS == 5614878 and N_WEIGHT_FLOATS_HFT == 5516574 and len(STATE_DICT_LAYOUT) == 165


def default_state_dict(seed=None, n_frame=N_FRAME):
    """Default-initialised parameters (what the reference gets for keys absent from the checkpoint): nn.Linear /
    nn.Conv2d kaiming-uniform(a=sqrt 5) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)), nn.Embedding N(0,1), LayerNorm (1, 0)."""
    g = torch.Generator()
    if seed is not None:
        g.manual_seed(int(seed))
    sd = {}
    lay = layout(n_frame)
    shapes = dict(lay)
    for key, shape in lay:
        if "layer_norm" in key:
            sd[key] = torch.ones(shape) if key.endswith("weight") else torch.zeros(shape)
        elif "pos_embedding" in key:
            sd[key] = torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:])) if key.endswith("weight") else None
            if fan_in is None:  # bias: fan_in of the matching weight
                wshape = shapes[key[:-4] + "weight"]
                fan_in = int(np.prod(wshape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def pack_state_dict(state_dict, strict=False, defaults=None, n_frame=N_FRAME):
    """Flattens a reference-format state_dict into the fp32 blob.  Unknown keys are ignored and missing keys take
    ``defaults`` when ``strict`` is False (reference behaviour, extractor.py:109); shape mismatches always raise."""
    lay = layout(n_frame)
    blob = np.empty(sum(int(np.prod(s)) for _, s in lay), dtype=np.float32)
    pos = 0
    missing = []
    for key, shape in lay:
        n = int(np.prod(shape))
        if key in state_dict:
            t = state_dict[key]
            t = t.detach().to("cpu", torch.float32).numpy() if isinstance(t, torch.Tensor) else np.asarray(t, np.float32)
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"size mismatch for {key}: checkpoint {tuple(t.shape)} vs model {tuple(shape)}")
            blob[pos : pos + n] = t.reshape(-1)
        else:
            missing.append(key)
            if strict:
                raise KeyError(f"missing key in state_dict: {key}")
            if defaults is None:
                defaults = default_state_dict(n_frame=n_frame)
            blob[pos : pos + n] = defaults[key].numpy().reshape(-1)
        pos += n
    return blob, missing
