"""etude_b200 -- B200-native (sm_100a) Extract stage of Etude: the AMT-APC / hFT-Transformer audio -> piano-roll ->
notes extractor, behind the reference's own Python class API.

    from etude_b200 import AMTAPC_Extractor, ExtractorConfig
    ex = AMTAPC_Extractor(ExtractorConfig(), "checkpoints/extractor/latest.pth", device="cuda")
    ex.extract("song.wav", "extract.json")

All arithmetic runs in ``libetude_b200.so`` (hand-written CUDA: tcgen05/TMEM GEMM + attention fed by TMA, fused
log-mel front-end, device note decoding) through the C ABI of ``include/etude_b200.h``.  No CPU fallback.
"""
from .config import ExtractorConfig  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require torch.cuda or the built library
    if name in ("AMTAPC_Extractor", "_load_model"):
        from . import extractor
        return getattr(extractor, name)
    if name in ("Model_SPEC2MIDI", "Encoder_SPEC2MIDI", "Decoder_SPEC2MIDI", "_Spec2MIDI"):
        from . import model
        return getattr(model, name)
    if name in ("HFT_Transformer", "HFTConfig"):
        from . import hft
        return getattr(hft, name)
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)
