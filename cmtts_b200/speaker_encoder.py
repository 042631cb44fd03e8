"""DeepSpeaker speaker encoder for the zero-shot path.

Mirrors the reference's interface around its TF-Keras model (no TensorFlow / h5py / librosa / python_speech_features in
this image, and none needed):

    build_model(ckpt_path)                                  deepspeaker/embedding.py:8-11
    predict_embedding(model, audio, sr, win_length, cuda)   deepspeaker/embedding.py:13-27   -> (1, 512) numpy
    PreDefinedEmbedder(config)(audio)                       speakerembedder/speaker_embedder.py:17-53
    get_deep_speaker_emb(filepath, batch_size, device)      call site synthesize_zeroshot_lj.py:93-97 -> (batch_size, 512)

Host side (numpy, as in the reference): `read_mfcc` = energy-percentile trim + python_speech_features.fbank (third-party,
not vendored in the reference and not pinned in its requirements.txt; its published algorithm — pre-emphasis 0.97,
25 ms / 10 ms rectangular frames, |rfft|^2 / nfft, triangular mel filters on floor((nfft + 1) f / sr) bins — is restated
here) + per-frame normalisation (audio_ds.py:33-44, :128-141), and `sample_from_mfcc` (batcher.py:23-29).
Device side: the ResCNN forward pass is `cmtts_rescnn_forward` of the C ABI (csrc/rescnn.cu), weights read from the
reference's Keras HDF5 checkpoint by `h5lite.H5File` and BatchNormalization folded on the host in fp64.  There is no CPU
fallback: the model needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import random
from decimal import ROUND_HALF_UP, Decimal
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib, ops
from .h5lite import H5File

# deepspeaker/constants.py:3-19
SAMPLE_RATE = 22050
WIN_LENGTH = 1024
NUM_FRAMES = 160
NUM_FBANKS = 64
BN_EPS = 1e-3                      # Keras BatchNormalization default (stored in the checkpoint's model_config)
DEFAULT_CKPT = "./deepspeaker/pretrained_models/ResCNN_triplet_training_checkpoint_265.h5"   # speaker_embedder.py:33


# ------------------------------------------------------------------------------------------------------------------
# host-side feature extraction
# ------------------------------------------------------------------------------------------------------------------
def calculate_nfft(samplerate: float, winlen: float) -> int:
    """audio_ds.py:17-30: smallest power of two >= the window length in samples."""
    n, nfft = winlen * samplerate, 1
    while nfft < n:
        nfft *= 2
    return nfft


def _round_half_up(x: float) -> int:
    return int(Decimal(x).quantize(Decimal("1"), rounding=ROUND_HALF_UP))


def mel_filterbank(nfilt: int, nfft: int, samplerate: float, lowfreq: float = 0.0, highfreq: Optional[float] = None) -> np.ndarray:
    """python_speech_features.get_filterbanks: (nfilt, nfft // 2 + 1) triangular filters, HTK mel scale."""
    highfreq = highfreq or samplerate / 2
    hz2mel = lambda hz: 2595.0 * np.log10(1.0 + hz / 700.0)
    mel2hz = lambda mel: 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    melpoints = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    bins = np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)
    fb = np.zeros((nfilt, nfft // 2 + 1))
    for j in range(nfilt):
        for i in range(int(bins[j]), int(bins[j + 1])):
            fb[j, i] = (i - bins[j]) / (bins[j + 1] - bins[j])
        for i in range(int(bins[j + 1]), int(bins[j + 2])):
            fb[j, i] = (bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])
    return fb


def fbank(signal: np.ndarray, samplerate: float, nfilt: int, nfft: int, winlen: float = 0.025, winstep: float = 0.01,
          preemph: float = 0.97) -> np.ndarray:
    """python_speech_features.fbank (filter-bank ENERGIES, no log; rectangular window), float64."""
    signal = np.asarray(signal, dtype=np.float64)
    signal = np.append(signal[0], signal[1:] - preemph * signal[:-1])
    flen, fstep = _round_half_up(winlen * samplerate), _round_half_up(winstep * samplerate)
    slen = len(signal)
    nframes = 1 if slen <= flen else 1 + int(math.ceil((1.0 * slen - flen) / fstep))
    padlen = (nframes - 1) * fstep + flen
    padded = np.concatenate((signal, np.zeros(padlen - slen)))
    idx = np.arange(flen)[None, :] + (np.arange(nframes) * fstep)[:, None]
    frames = padded[idx]
    pspec = (1.0 / nfft) * np.square(np.abs(np.fft.rfft(frames, nfft)))
    feat = pspec @ mel_filterbank(nfilt, nfft, samplerate).T
    return np.where(feat == 0, np.finfo(float).eps, feat)


def normalize_frames(m: np.ndarray, epsilon: float = 1e-12) -> np.ndarray:
    """audio_ds.py:140-141: every frame to zero mean / unit variance over its filters."""
    return np.stack([(v - np.mean(v)) / max(np.std(v), epsilon) for v in m])


def read_mfcc(audio: np.ndarray, sample_rate: int, win_length: int) -> np.ndarray:
    """audio_ds.py:33-44 (+ mfcc_fbank :128-137): trim to the span of samples louder than the 95th percentile, 64 mel
    filter-bank energies per 25 ms frame, per-frame normalisation -> (frames, 64) float32."""
    audio = np.asarray(audio)
    energy = np.abs(audio)
    offsets = np.where(energy > np.percentile(energy, 95))[0]
    if offsets.size == 0:
        raise ValueError("read_mfcc: silent audio (no sample above the 95th percentile)")
    voice = audio[offsets[0]:offsets[-1]]
    nfft = calculate_nfft(sample_rate, win_length / sample_rate)
    feats = normalize_frames(fbank(voice, sample_rate, NUM_FBANKS, nfft))
    return np.array(feats, dtype=np.float32)


def pad_mfcc(mfcc: np.ndarray, max_length: int) -> np.ndarray:
    if len(mfcc) < max_length:
        mfcc = np.vstack((mfcc, np.zeros((max_length - len(mfcc), mfcc.shape[1]), mfcc.dtype)))
    return mfcc


def sample_from_mfcc(mfcc: np.ndarray, max_length: int, offset: Optional[int] = None) -> np.ndarray:
    """batcher.py:23-29: a random `max_length`-frame window (the reference draws it with random.choice; pass `offset` to
    fix it), zero-padded if the utterance is shorter; trailing channel axis added."""
    if mfcc.shape[0] >= max_length:
        r = random.choice(range(0, len(mfcc) - max_length + 1)) if offset is None else int(offset)
        s = mfcc[r:r + max_length]
    else:
        s = pad_mfcc(mfcc, max_length)
    return np.expand_dims(s, axis=-1)


def load_audio(path: str, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """WAV file -> mono float32 in [-1, 1) at `sample_rate` (the reference uses librosa.load; files already at the target
    rate — LJSpeech at 22050 Hz — decode identically, others are resampled with a polyphase filter instead of librosa's)."""
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.dtype.kind == "i":
        data = data.astype(np.float32) / float(2 ** (8 * data.dtype.itemsize - 1))
    elif data.dtype.kind == "u":
        data = (data.astype(np.float32) - 128.0) / 128.0
    data = data.astype(np.float32)
    if data.ndim > 1:
        data = data.mean(axis=1)
    if sr != sample_rate:
        from scipy.signal import resample_poly
        g = math.gcd(int(sr), int(sample_rate))
        data = resample_poly(data, sample_rate // g, sr // g).astype(np.float32)
    return data


# ------------------------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------------------------
STAGE_FILTERS = (64, 128, 256, 512)


def conv_layer_names():
    """Keras layer names of the 28 convs in network order (conv_models.py:83-131)."""
    names = []
    for stage, f in enumerate(STAGE_FILTERS, start=1):
        names.append(f"conv{f}-s")
        for blk in range(3):
            names += [f"res{stage}_{blk}_branch_2a", f"res{stage}_{blk}_branch_2b"]
    return names


def fold_weights(w: Dict[str, np.ndarray]):
    """Keras weights {'<layer>/kernel:0', ...} -> the table of cmtts_rescnn_forward: per conv (kernel HWIO fp32, scale,
    shift) with BatchNormalization and the conv bias folded in fp64, then the dense kernel and bias."""
    out = []
    for name in conv_layer_names():
        k = np.asarray(w[f"{name}/kernel:0"], dtype=np.float32)
        b = np.asarray(w[f"{name}/bias:0"], dtype=np.float64)
        g, beta = (np.asarray(w[f"{name}_bn/{t}:0"], dtype=np.float64) for t in ("gamma", "beta"))
        mean, var = (np.asarray(w[f"{name}_bn/{t}:0"], dtype=np.float64) for t in ("moving_mean", "moving_variance"))
        scale = g / np.sqrt(var + BN_EPS)
        shift = beta + (b - mean) * scale
        out += [np.ascontiguousarray(k), scale.astype(np.float32), shift.astype(np.float32)]
    out += [np.ascontiguousarray(np.asarray(w["affine/kernel:0"], dtype=np.float32)),
            np.asarray(w["affine/bias:0"], dtype=np.float32)]
    return out


def read_keras_weights(path: str) -> Dict[str, np.ndarray]:
    """{'<layer>/<weight>:0': array} from a Keras HDF5 checkpoint (group model_weights/<layer>/<layer>/<weight>:0)."""
    f = H5File(path)
    root = "model_weights" if "model_weights" in f.keys("/") else "/"
    res = {}
    for p, a in f.datasets(root).items():
        parts = p.split("/")
        res[f"{parts[0]}/{parts[-1]}"] = a
    return res


class _Packed:
    def __init__(self, tensors):
        self.tensors = tensors
        self.ptrs = _lib.pointer_table(tensors)
        self.cfg = (C.c_int32 * 3)(NUM_FBANKS, STAGE_FILTERS[0], 512)


class DeepSpeakerModel:
    """conv_models.py:22-66 (inference configuration: include_softmax=False).  `.m` is the object itself, so the
    reference's `model.m.predict(x)` / `model.m.load_weights(path, by_name=True)` calls read the same."""

    def __init__(self, device: Optional[str] = None):
        self.device = torch.device(device or "cuda")
        self.lib = _lib.load()
        self.packed: Optional[_Packed] = None
        self.handle = ops.register(self)
        self._ws: Optional[torch.Tensor] = None
        self.m = self

    def __del__(self):
        try:
            ops.release(self.handle)
        except Exception:
            pass

    def set_keras_weights(self, w: Dict[str, np.ndarray]) -> "DeepSpeakerModel":
        if self.device.type != "cuda":
            raise _lib.CmttsError("DeepSpeakerModel needs a CUDA device (there is no CPU fallback)")
        table = fold_weights(w)
        expect_cin = 1
        for i, name in enumerate(conv_layer_names()):
            k = table[3 * i]
            f = STAGE_FILTERS[i // 7]
            ks = 5 if i % 7 == 0 else 3
            if k.shape != (ks, ks, expect_cin, f):
                raise ValueError(f"{name}: kernel {k.shape}, expected {(ks, ks, expect_cin, f)}")
            expect_cin = f
        if table[-2].shape != (4 * 512, 512):
            raise ValueError(f"affine: kernel {table[-2].shape}, expected (2048, 512)")
        self.packed = _Packed([torch.from_numpy(np.ascontiguousarray(t)).to(self.device) for t in table])
        return self

    def load_weights(self, path: str, by_name: bool = True) -> "DeepSpeakerModel":
        return self.set_keras_weights(read_keras_weights(path))

    def predict(self, x) -> np.ndarray:
        """(B, T, 64, 1) or (B, T, 64) normalised filter-bank frames -> (B, 512) numpy (keras Model.predict)."""
        return self.predict_tensor(torch.as_tensor(np.asarray(x), dtype=torch.float32)).cpu().numpy()

    def predict_tensor(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 4:
            if x.shape[-1] != 1:
                raise ValueError("DeepSpeakerModel: the input has one channel")
            x = x[..., 0]
        if x.dim() != 3 or x.shape[2] != NUM_FBANKS:
            raise ValueError(f"DeepSpeakerModel: input must be (B, T, {NUM_FBANKS}[, 1])")
        return torch.ops.cmtts_b200.rescnn_forward(self.handle, x.to(self.device, torch.float32).contiguous())


def build_model(ckpt_path: str, device: Optional[str] = None) -> DeepSpeakerModel:
    """deepspeaker/embedding.py:8-11."""
    model = DeepSpeakerModel(device)
    model.m.load_weights(ckpt_path, by_name=True)
    return model


def predict_embedding(model: DeepSpeakerModel, audio: np.ndarray, sr: int = SAMPLE_RATE, win_length: int = WIN_LENGTH,
                      cuda: bool = True, offset: Optional[int] = None) -> np.ndarray:
    """deepspeaker/embedding.py:13-27 -> (1, 512).  `cuda=False` (TensorFlow on the CPU in the reference) is refused: this
    implementation has no CPU path.  `offset` fixes the frame window the reference draws at random."""
    if not cuda:
        raise _lib.CmttsError("predict_embedding: no CPU path (cuda=False)")
    mfcc = sample_from_mfcc(read_mfcc(audio, sr, win_length), NUM_FRAMES, offset)
    return model.m.predict(np.expand_dims(mfcc, axis=0))


class PreDefinedEmbedder:
    """speakerembedder/speaker_embedder.py:17-53 ("DeepSpeaker" only; the GE2E LSTM encoder is not on any shipped
    zero-shot script's path)."""

    def __init__(self, config, ckpt_path: Optional[str] = None, device: Optional[str] = None):
        self.sampling_rate = config.sampling_rate
        self.win_length = config.win_length
        self.embedder_type = config.speaker_embedder
        self.embedder_cuda = getattr(config, "speaker_embedder_cuda", True)
        self.config = config
        if self.embedder_type != "DeepSpeaker":
            raise NotImplementedError(f"speaker embedder {self.embedder_type!r}: only 'DeepSpeaker'")
        self.embedder = build_model(ckpt_path or DEFAULT_CKPT, device)

    def forward(self, audio):
        return predict_embedding(self.embedder, audio, self.sampling_rate, self.win_length, self.embedder_cuda)

    __call__ = forward


_EMBEDDERS: Dict[str, DeepSpeakerModel] = {}


def get_deep_speaker_emb(filepath: str, batch_size: int, device, ckpt_path: Optional[str] = None,
                         sampling_rate: int = SAMPLE_RATE, win_length: int = WIN_LENGTH, offset: Optional[int] = None) -> torch.Tensor:
    """The zero-shot scripts' call (synthesize_zeroshot_lj.py:93-97: `from speakerembedder import get_deep_speaker_emb`; the
    shipped speakerembedder/__init__.py does not define it — this is what the call site implies): the DeepSpeaker embedding
    of one reference recording, repeated for every utterance of the batch -> (batch_size, 512) float32 on `device`.
    The model is loaded once per checkpoint path and device."""
    ckpt = ckpt_path or DEFAULT_CKPT
    dev = torch.device(device)
    key = f"{os.path.abspath(ckpt)}@{dev}"
    if key not in _EMBEDDERS:
        _EMBEDDERS[key] = build_model(ckpt, str(dev))
    audio = load_audio(filepath, sampling_rate)
    mfcc = sample_from_mfcc(read_mfcc(audio, sampling_rate, win_length), NUM_FRAMES, offset)
    emb = _EMBEDDERS[key].predict_tensor(torch.from_numpy(np.expand_dims(mfcc, 0)))
    return emb.expand(int(batch_size), -1).contiguous()
