"""cmtts_b200 — B200-native (sm_100a) implementation of the CM-TTS inference hot path.

FastSpeech2 FFT encoder + variance adaptor -> consistency-model denoiser (T in {1,2,4}) ->
HiFi-GAN generator, behind the reference's call protocols:

    cmtts_b200.model      CMTotalTTS, KarrasDenoiser, create_model_and_diffusion_tts
    cmtts_b200.sampler    karras_sample_tts, get_sigmas_karras
    cmtts_b200.vocoder    Generator, get_vocoder, vocoder_infer
    cmtts_b200.synthesize CMTotalTTSSynthesize, Pipeline

All arithmetic runs in cmtts_b200/lib/libcmtts_b200.so (hand-written CUDA behind the C ABI of
include/cmtts_b200.h); importing this package does not need a GPU, running it does.
"""
from .config import HifiGanSpec, ModelSpec  # noqa: F401

__version__ = "0.1.0"
