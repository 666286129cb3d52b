"""ExtractorConfig defaults, restated from the reference's pydantic schema (etude/config/schema.py:68-131).

The extractor accepts either these dataclasses or the reference's own ``load_config().extractor`` object (same
attribute tree).  The CUDA build is specialised to the default shapes; ``validate`` rejects anything else with a
clear error -- never a fallback.
"""
from dataclasses import dataclass, field


@dataclass
class ExtractorFeatureConfig:
    sr: int = 16000
    hop_sample: int = 256
    mel_bins: int = 256
    n_bins: int = 256
    fft_bins: int = 2048
    window_length: int = 2048
    log_offset: float = 1e-8
    window: str = "hann"
    pad_mode: str = "constant"


@dataclass
class ExtractorInputConfig:
    margin_b: int = 32
    margin_f: int = 32
    num_frame: int = 512
    min_value: float = -18.0


@dataclass
class ExtractorMidiConfig:
    note_min: int = 21
    note_max: int = 108
    num_note: int = 88
    num_velocity: int = 128


@dataclass
class ExtractorModelConfig:
    cnn_channel: int = 4
    cnn_kernel: int = 5
    dropout: float = 0.1
    transformer_hid_dim: int = 256
    transformer_pf_dim: int = 512
    encoder_n_head: int = 4
    encoder_n_layer: int = 3
    decoder_n_head: int = 4
    decoder_n_layer: int = 3
    sv_dim: int = 24


@dataclass
class ExtractorInferConfig:
    onset_threshold: float = 0.5
    offset_threshold: float = 1.0
    frame_threshold: float = 0.5
    min_duration: float = 0.08


@dataclass
class ExtractorConfig:
    feature: ExtractorFeatureConfig = field(default_factory=ExtractorFeatureConfig)
    input: ExtractorInputConfig = field(default_factory=ExtractorInputConfig)
    midi: ExtractorMidiConfig = field(default_factory=ExtractorMidiConfig)
    model: ExtractorModelConfig = field(default_factory=ExtractorModelConfig)
    infer: ExtractorInferConfig = field(default_factory=ExtractorInferConfig)


_SHAPE_FIELDS = [
    ("feature", ["sr", "hop_sample", "mel_bins", "n_bins", "fft_bins", "window_length", "log_offset"]),
    ("input", ["margin_b", "margin_f", "num_frame", "min_value"]),
    ("midi", ["num_note", "num_velocity"]),
    ("model", ["cnn_channel", "cnn_kernel", "transformer_hid_dim", "transformer_pf_dim", "encoder_n_head",
               "encoder_n_layer", "decoder_n_head", "decoder_n_layer"]),
]


def validate(config):
    """Raises ValueError unless every shape-defining field equals the default the kernels are compiled for."""
    ref = ExtractorConfig()
    for section, names in _SHAPE_FIELDS:
        for n in names:
            got, want = getattr(getattr(config, section), n), getattr(getattr(ref, section), n)
            if got != want:
                raise ValueError(f"etude_b200 is compiled for extractor.{section}.{n} = {want!r}; got {got!r}. "
                                 "Other shapes are not supported (there is no fallback path).")
    return config
