"""TEST INFRASTRUCTURE — CPU restatement (numpy, float64 accumulation) of the reference's DeepSpeaker ResCNN forward pass
for the zero-shot path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

PARITY UNPINNED: the reference model is TF-Keras (deepspeaker/conv_models.py) and this image has no TensorFlow, so the
restatement cannot be run against the reference itself.  What it is anchored on instead:
  * the layer graph, names, shapes and hyper-parameters are read from the `model_config` JSON that Keras stored inside the
    reference's own checkpoint (ResCNN_triplet_training_checkpoint_265.h5: kernel sizes, strides, padding "same",
    BatchNormalization epsilon 0.001 / axis 3, Dense 512) — tests/test_speaker_encoder.py checks this file against it;
  * Keras / TensorFlow layer semantics as published: Conv2D "same" padding with stride s pads
    max((ceil(n / s) - 1) s + k - n, 0) in total, the odd element AFTER; BatchNormalization at inference is
    gamma (x - moving_mean) / sqrt(moving_var + eps) + beta; K.l2_normalize(x) = x / sqrt(max(sum x^2, 1e-12)).

Each function cites the reference lines it follows.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

STAGE_FILTERS = (64, 128, 256, 512)
BN_EPS = 1e-3


def _same_pads(n: int, k: int, s: int):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def conv2d_same(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray, stride: int) -> np.ndarray:
    """keras Conv2D(padding='same', data_format channels_last): x (B, H, W, Cin), kernel (kh, kw, Cin, Cout)."""
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = kernel.shape
    Ho, pt, pb = _same_pads(H, kh, stride)
    Wo, pl, pr = _same_pads(W, kw, stride)
    xp = np.pad(x.astype(np.float64), ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    out = np.zeros((B, Ho, Wo, Cout))
    k64 = kernel.astype(np.float64)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, i:i + (Ho - 1) * stride + 1:stride, j:j + (Wo - 1) * stride + 1:stride, :]
            out += patch @ k64[i, j]
    return out + bias.astype(np.float64)


def batchnorm(x, w: Dict[str, np.ndarray], name: str) -> np.ndarray:
    g, b = w[f"{name}/gamma:0"].astype(np.float64), w[f"{name}/beta:0"].astype(np.float64)
    m, v = w[f"{name}/moving_mean:0"].astype(np.float64), w[f"{name}/moving_variance:0"].astype(np.float64)
    return g * (x - m) / np.sqrt(v + BN_EPS) + b


def clipped_relu(x):
    """conv_models.py:78-81: K.minimum(K.maximum(y, 0), 20)."""
    return np.minimum(np.maximum(x, 0.0), 20.0)


def identity_block(x, w, stage: int, block: int):
    """conv_models.py:83-108 (note the order: conv, BN, clip, conv, BN, clip, THEN add, clip)."""
    base = f"res{stage}_{block}_branch"
    y = conv2d_same(x, w[f"{base}_2a/kernel:0"], w[f"{base}_2a/bias:0"], 1)
    y = clipped_relu(batchnorm(y, w, f"{base}_2a_bn"))
    y = conv2d_same(y, w[f"{base}_2b/kernel:0"], w[f"{base}_2b/bias:0"], 1)
    y = clipped_relu(batchnorm(y, w, f"{base}_2b_bn"))
    return clipped_relu(y + x)


def conv_and_res_block(x, w, filters: int, stage: int):
    """conv_models.py:110-126."""
    name = f"conv{filters}-s"
    o = conv2d_same(x, w[f"{name}/kernel:0"], w[f"{name}/bias:0"], 2)
    o = clipped_relu(batchnorm(o, w, f"{name}_bn"))
    for i in range(3):
        o = identity_block(o, w, stage, i)
    return o


def rescnn_forward(x: np.ndarray, w: Dict[str, np.ndarray], return_stages: bool = False):
    """conv_models.py:44-66, :128-133: x (B, T, 64, 1) -> (B, 512) L2-normalised embedding (float64)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 3:
        x = x[..., None]
    stages = []
    for stage, f in enumerate(STAGE_FILTERS, start=1):
        x = conv_and_res_block(x, w, f, stage)
        stages.append(x)
    B = x.shape[0]
    x = x.reshape(B, -1, 2048)                     # Reshape((-1, 2048))
    x = x.mean(axis=1)                             # Lambda K.mean(axis=1), name 'average'
    x = x @ w["affine/kernel:0"].astype(np.float64) + w["affine/bias:0"].astype(np.float64)
    x = x / np.sqrt(np.maximum((x * x).sum(axis=1, keepdims=True), 1e-12))   # K.l2_normalize(axis=1)
    return (x, stages) if return_stages else x


def synthetic_keras_weights(seed: int = 0, scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Random weights in the checkpoint's naming / shapes (for tests on boxes without the 97 MB file)."""
    from cmtts_b200.synthetic import make_deepspeaker_weights
    return make_deepspeaker_weights(seed, scale)
