"""Host-side mirror of the reference model classes (etude/models/amt_apc.py) over the CUDA engine.

``Model_SPEC2MIDI(encoder, decoder).forward(input_spec[B,256,576])`` returns the same 9-tuple as the reference
(amt_apc.py:29-49); ``Encoder_SPEC2MIDI`` / ``Decoder_SPEC2MIDI`` take the reference constructor arguments
(amt_apc.py:56, 127) and expose parameters under the reference's state_dict keys, so
``checkpoints/extractor/latest.pth`` loads unchanged.  The arithmetic runs in libetude_b200.so; these modules only
hold parameters.  Inference only (eval mode, no autograd), default ExtractorConfig shapes only.
"""
import torch
import torch.nn as nn

from . import engine as _engine
from .weights import STATE_DICT_LAYOUT, default_state_dict, pack_state_dict


class _ParamTree(nn.Module):
    """Parameter container reproducing one sub-tree ("encoder." / "decoder.") of the reference state_dict."""

    def __init__(self, prefix, init):
        super().__init__()
        for key, shape in STATE_DICT_LAYOUT:
            if not key.startswith(prefix + "."):
                continue
            parts = key[len(prefix) + 1:].split(".")
            mod = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, nn.Module())
                mod = mod._modules[p]
            mod.register_parameter(parts[-1], nn.Parameter(init[key].clone(), requires_grad=False))


def _check(name, got, want):
    if got != want:
        raise ValueError(f"etude_b200 is compiled for {name}={want}; got {got} (no fallback path for other shapes)")


class Encoder_SPEC2MIDI(_ParamTree):
    def __init__(self, n_margin=32, n_frame=512, n_bin=256, cnn_channel=4, cnn_kernel=5, hid_dim=256, n_layers=3, n_heads=4,
                 pf_dim=512, dropout=0.1, device=None):
        for n, g, w in [("n_margin", n_margin, 32), ("n_frame", n_frame, 512), ("n_bin", n_bin, 256), ("cnn_channel", cnn_channel, 4),
                        ("cnn_kernel", cnn_kernel, 5), ("hid_dim", hid_dim, 256), ("n_layers", n_layers, 3), ("n_heads", n_heads, 4),
                        ("pf_dim", pf_dim, 512)]:
            _check(n, g, w)
        super().__init__("encoder", default_state_dict())
        self.hid_dim, self.n_frame, self.n_bin, self.device = hid_dim, n_frame, n_bin, device


class Decoder_SPEC2MIDI(_ParamTree):
    def __init__(self, n_frame=512, n_bin=256, n_note=88, n_velocity=128, hid_dim=256, n_layers=3, n_heads=4, pf_dim=512,
                 dropout=0.1, device=None):
        for n, g, w in [("n_frame", n_frame, 512), ("n_bin", n_bin, 256), ("n_note", n_note, 88), ("n_velocity", n_velocity, 128),
                        ("hid_dim", hid_dim, 256), ("n_layers", n_layers, 3), ("n_heads", n_heads, 4), ("pf_dim", pf_dim, 512)]:
            _check(n, g, w)
        super().__init__("decoder", default_state_dict())
        self.hid_dim, self.n_frame, self.n_note, self.device = hid_dim, n_frame, n_note, device


class Model_SPEC2MIDI(nn.Module):
    """Reference: etude/models/amt_apc.py:23-49.  Parameters live under ``encoder_spec2midi`` / ``decoder_spec2midi``."""

    _enc_name, _dec_name = "encoder_spec2midi", "decoder_spec2midi"

    def __init__(self, encoder, decoder, max_windows=8):
        super().__init__()
        setattr(self, self._enc_name, encoder)
        setattr(self, self._dec_name, decoder)
        self._engine = None
        self._engine_device = None
        self._max_windows = max_windows

    # --- engine management: weights are packed once, on first use after (re)loading
    def _flat_state_dict(self):
        sd = {}
        for k, v in getattr(self, self._enc_name).state_dict().items():
            sd["encoder." + k] = v
        for k, v in getattr(self, self._dec_name).state_dict().items():
            sd["decoder." + k] = v
        return sd

    def load_state_dict(self, state_dict, strict=True, assign=False):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def engine(self, device=None):
        if device is None:
            device = next(self.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("etude_b200 models run on CUDA devices only (move the model with .to('cuda')); no CPU fallback")
        if self._engine is None or self._engine_device != device:
            blob, _ = pack_state_dict(self._flat_state_dict(), strict=True)
            self._engine = _engine.Engine(blob, device, max_windows=self._max_windows)
            self._engine_device = device
        return self._engine

    @torch.no_grad()
    def forward(self, input_spec):
        if input_spec.dim() != 3 or input_spec.shape[1] != 256 or input_spec.shape[2] != 576:
            raise ValueError(f"input_spec must be [B, 256, 576], got {tuple(input_spec.shape)}")
        eng = self.engine(input_spec.device if input_spec.is_cuda else None)
        dev = eng.device
        b = input_spec.shape[0]
        feat = input_spec.to(dev, torch.float32).transpose(1, 2).contiguous().reshape(b * 576, 256)
        rolls_a = eng.alloc_rolls(b * 512, dev)
        rolls_b = eng.alloc_rolls(b * 512, dev)
        vel_a = torch.empty((b, 512, 88, 128), dtype=torch.float32, device=dev)
        vel_b = torch.empty((b, 512, 88, 128), dtype=torch.float32, device=dev)
        att = torch.empty((b, 512, 4, 88, 256), dtype=torch.float32, device=dev)
        eng.forward_windows(feat, [i * 576 for i in range(b)], [i * 512 for i in range(b)], rolls_b, rolls_a, vel_a, vel_b, att)
        r = lambda t: t.reshape(b, 512, 88)
        return (r(rolls_a[0]), r(rolls_a[1]), r(rolls_a[2]), vel_a, att, r(rolls_b[0]), r(rolls_b[1]), r(rolls_b[2]), vel_b)


class _Spec2MIDI(Model_SPEC2MIDI):
    """Reference: etude/data/extractor.py:34-75 (style-vector branch disabled: sv_dim = 0, extractor.py:107).
    Parameters live under ``encoder`` / ``decoder`` (state_dict keys ``encoder.*`` / ``decoder.*``)."""

    _enc_name, _dec_name = "encoder", "decoder"

    def __init__(self, encoder, decoder, sv_dim=0, max_windows=32):
        if sv_dim:
            raise ValueError("the style-vector branch (sv_dim != 0) is disabled in the reference extractor and not built here")
        super().__init__(encoder, decoder, max_windows=max_windows)
        self.sv_dim = sv_dim

    def forward(self, x, sv=None):
        return super().forward(x)
