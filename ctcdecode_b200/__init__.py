"""ctcdecode_b200: B200-native CTC prefix beam search behind the parlance/ctcdecode Python interface.

    from ctcdecode_b200 import CTCBeamDecoder
    beam_results, beam_scores, timesteps, out_lens = CTCBeamDecoder(labels).decode(probs)

Only the beam-search hot path is here; the CUDA library must be built
(`python -m ctcdecode_b200.build`) and a B200 present -- there is no CPU fallback.
"""
from .decoder import CTCBeamDecoder, DecoderState, OnlineCTCBeamDecoder, convert_to_string  # noqa: F401

__all__ = ["CTCBeamDecoder", "OnlineCTCBeamDecoder", "DecoderState", "convert_to_string"]
