"""Public surface of the B200-native FastSpeech2-align forward (imported as `smart_nar_fast_tts_b200`)."""
from .capi import Fs2Error, Fs2Library, PREC_BF16, PREC_BF16X3, PREC_F16X2, PREC_FP32, load_library  # noqa: F401
from .model import FastSpeech2Align, dims_from_configs  # noqa: F401
from .sharding import ShardedSynthesizer, shard_bounds  # noqa: F401
from .streamed import StreamedSynthesizer  # noqa: F401
from .operators import GaussianUpsampling, LengthRegulator, get_mask_from_lengths  # noqa: F401
from . import operators, pipeline, synthetic  # noqa: F401

__all__ = ["FastSpeech2Align", "dims_from_configs", "Fs2Error", "Fs2Library", "load_library", "PREC_FP32", "PREC_BF16", "PREC_BF16X3", "PREC_F16X2",
           "ShardedSynthesizer", "StreamedSynthesizer", "shard_bounds", "synthetic", "pipeline", "operators",
           "LengthRegulator", "GaussianUpsampling", "get_mask_from_lengths"]
