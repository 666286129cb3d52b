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
from .weights import default_state_dict, layout, pack_state_dict


class _ParamTree(nn.Module):
    """Parameter container reproducing one sub-tree ("encoder." / "decoder.") of the reference state_dict."""

    def __init__(self, prefix, init, n_frame=512):
        super().__init__()
        for key, shape in layout(n_frame):
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


def _check_frame(n_frame):
    if n_frame not in (512, 128):
        raise ValueError(f"etude_b200 is compiled for n_frame=512 (ExtractorConfig) or 128 (HFTConfig); got {n_frame}")


class Encoder_SPEC2MIDI(_ParamTree):
    def __init__(self, n_margin=32, n_frame=512, n_bin=256, cnn_channel=4, cnn_kernel=5, hid_dim=256, n_layers=3, n_heads=4,
                 pf_dim=512, dropout=0.1, device=None):
        _check_frame(n_frame)
        for n, g, w in [("n_margin", n_margin, 32), ("n_bin", n_bin, 256), ("cnn_channel", cnn_channel, 4),
                        ("cnn_kernel", cnn_kernel, 5), ("hid_dim", hid_dim, 256), ("n_layers", n_layers, 3), ("n_heads", n_heads, 4),
                        ("pf_dim", pf_dim, 512)]:
            _check(n, g, w)
        super().__init__("encoder", default_state_dict(n_frame=n_frame), n_frame)
        self.hid_dim, self.n_frame, self.n_bin, self.device = hid_dim, n_frame, n_bin, device


class Decoder_SPEC2MIDI(_ParamTree):
    def __init__(self, n_frame=512, n_bin=256, n_note=88, n_velocity=128, hid_dim=256, n_layers=3, n_heads=4, pf_dim=512,
                 dropout=0.1, device=None):
        _check_frame(n_frame)
        for n, g, w in [("n_bin", n_bin, 256), ("n_note", n_note, 88), ("n_velocity", n_velocity, 128),
                        ("hid_dim", hid_dim, 256), ("n_layers", n_layers, 3), ("n_heads", n_heads, 4), ("pf_dim", pf_dim, 512)]:
            _check(n, g, w)
        super().__init__("decoder", default_state_dict(n_frame=n_frame), n_frame)
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

    @property
    def n_frame(self):
        return int(getattr(self, self._dec_name).n_frame)

    def engine(self, device=None):
        if device is None:
            device = next(self.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("etude_b200 models run on CUDA devices only (move the model with .to('cuda')); no CPU fallback")
        if self._engine is None or self._engine_device != device:
            if getattr(self, self._enc_name).n_frame != self.n_frame:
                raise ValueError("encoder and decoder were built for different n_frame")
            blob, _ = pack_state_dict(self._flat_state_dict(), strict=True, n_frame=self.n_frame)
            self._engine = _engine.Engine(blob, device, max_windows=self._max_windows)
            self._engine_device = device
        return self._engine

    def _rows(self, input_spec):
        """[B, 256, n_frame + 64] -> (engine, padded feature rows [B * (n_frame + 64), 256] on the device)."""
        f = self.n_frame
        if input_spec.dim() != 3 or input_spec.shape[1] != 256 or input_spec.shape[2] != f + 64:
            raise ValueError(f"input_spec must be [B, 256, {f + 64}], got {tuple(input_spec.shape)}")
        eng = self.engine(input_spec.device if input_spec.is_cuda else None)
        b = input_spec.shape[0]
        return eng, input_spec.to(eng.device, torch.float32).transpose(1, 2).contiguous().reshape(b * (f + 64), 256)

    def _decode_into(self, eng, b, feat=None, enc=None):
        f, dev = self.n_frame, eng.device
        rolls_a = eng.alloc_rolls(b * f, dev)
        rolls_b = eng.alloc_rolls(b * f, dev)
        vel_a = torch.empty((b, f, 88, 128), dtype=torch.float32, device=dev)
        vel_b = torch.empty((b, f, 88, 128), dtype=torch.float32, device=dev)
        att = torch.empty((b, f, 4, 88, 256), dtype=torch.float32, device=dev)
        eng.forward_windows(feat, [i * (f + 64) for i in range(b)], [i * f for i in range(b)], rolls_b, rolls_a, vel_a, vel_b, att,
                            enc_in=enc)
        r = lambda t: t.reshape(b, f, 88)
        return (r(rolls_a[0]), r(rolls_a[1]), r(rolls_a[2]), vel_a, att, r(rolls_b[0]), r(rolls_b[1]), r(rolls_b[2]), vel_b)

    @torch.no_grad()
    def forward(self, input_spec):
        eng, feat = self._rows(input_spec)
        return self._decode_into(eng, input_spec.shape[0], feat=feat)

    @torch.no_grad()
    def encode_spec(self, input_spec):
        """Encoder_SPEC2MIDI.forward (amt_apc.py:74-120): [B, 256, n_frame + 64] -> fp32 [B, n_frame, 256, 256]."""
        eng, feat = self._rows(input_spec)
        f = self.n_frame
        return eng.encode_windows(feat, [i * (f + 64) for i in range(input_spec.shape[0])])

    @torch.no_grad()
    def decode_spec(self, enc):
        """Decoder_SPEC2MIDI.forward (amt_apc.py:159-230) on an encoder output [B, n_frame, 256, 256] -> the 9-tuple."""
        f = self.n_frame
        if enc.dim() != 4 or tuple(enc.shape[1:]) != (f, 256, 256):
            raise ValueError(f"encoder output must be [B, {f}, 256, 256], got {tuple(enc.shape)}")
        eng = self.engine(enc.device if enc.is_cuda else None)
        h = enc.to(eng.device, torch.float32).contiguous()
        return self._decode_into(eng, h.shape[0], enc=h.reshape(-1, 256))


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
        """decode(encode(x)) (extractor.py:53-56) without the fp32 round trip of the encoder output -- bit-identical to it
        (activations are bf16 inside; tests/test_kernels_gpu.py::test_encode_decode_split)."""
        return super().forward(x)

    def encode(self, x, sv=None):
        """Reference: extractor.py:58-68 (the style-vector gate is compiled out: sv_dim = 0, so ``sv`` is ignored there too)."""
        return self.encode_spec(x)

    def decode(self, h):
        """Reference: extractor.py:70-75."""
        return self.decode_spec(h)
