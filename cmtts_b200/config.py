"""Hyper-parameters of the CM-TTS inference hot path.

The reference spreads these over three YAML files per dataset (config/<dataset>/{preprocess,
model,train}.yaml, read by utils/tools.py:25-33) plus `<preprocessed_path>/stats.json`
(model/modules.py:233-237).  `ModelSpec.from_reference_configs` consumes those dicts unchanged;
the three presets below restate the shipped values so the GPU box needs no YAML.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field, asdict
from typing import Optional, Tuple


@dataclass(frozen=True)
class HifiGanSpec:
    """hifigan/config.json:11-15 (HiFi-GAN V1)."""
    n_mels: int = 80
    upsample_rates: Tuple[int, ...] = (8, 8, 2, 2)
    upsample_kernel_sizes: Tuple[int, ...] = (16, 16, 4, 4)
    upsample_initial_channel: int = 512
    resblock_kernel_sizes: Tuple[int, ...] = (3, 7, 11)
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    lrelu_slope: float = 0.1          # hifigan/models.py:7
    final_lrelu_slope: float = 0.01   # F.leaky_relu default, hifigan/models.py:161

    @property
    def hop(self) -> int:
        h = 1
        for u in self.upsample_rates:
            h *= u
        return h


@dataclass(frozen=True)
class ModelSpec:
    name: str = "LJSpeech"
    # text encoder (config/*/model.yaml:1-12, text/symbols.py:21-29)
    vocab: int = 361
    hidden: int = 256
    enc_layers: int = 4
    enc_heads: int = 2
    ffn_kernel: int = 9
    ffn_act: str = "gelu"
    # variance adaptor (model.yaml:34-50)
    filter_size: int = 256
    dur_layers: int = 2
    dur_kernel: int = 3
    pred_layers: int = 2
    pred_kernel: int = 5
    cwt_hidden: int = 128
    cwt_std_scale: float = 0.8
    pitch_bins: int = 300
    energy_bins: int = 256
    use_uv: bool = True
    pitch_norm_eps: float = 1e-9
    energy_min: float = -1.5
    energy_max: float = 8.0
    # denoiser (model.yaml:14-25)
    n_mels: int = 80
    res_layers: int = 20
    res_channels: int = 256
    multi_speaker: bool = False
    ext_speaker_dim: int = 512
    # `speaker_embedder: none` (cmtts.py:27-37): an nn.Embedding(n_speaker, hidden) table indexed by the batch's speaker ids
    # instead of a Linear over external 512-d embeddings.  0 = external embeddings.  On the device the table is the
    # Linear's weight over one-hot rows (exact: one product with 1.0, the rest with 0.0), ext_speaker_dim = n_speakers
    # rounded up to a multiple of 16.
    n_speakers: int = 0
    # consistency model (train.yaml cm block; karras_diffusion.py / script_util.py:66-73)
    sigma_min: float = 0.002
    sigma_max: float = 80.0
    sigma_data: float = 0.5
    rho: float = 7.0
    # audio
    sampling_rate: int = 22050
    hop_length: int = 256
    max_wav_value: float = 32768.0
    max_seq_len: int = 1000
    vocoder_speaker: str = "LJSpeech"
    hifigan: HifiGanSpec = field(default_factory=HifiGanSpec)

    @property
    def cwt_out(self) -> int:
        return 11 if self.use_uv else 10

    @property
    def head_dim(self) -> int:
        return self.hidden // self.enc_heads

    def to_json(self) -> str:
        return json.dumps(asdict(self))

    @staticmethod
    def preset(dataset: str) -> "ModelSpec":
        if dataset == "LJSpeech":
            return ModelSpec(name="LJSpeech")
        if dataset == "VCTK":
            return ModelSpec(name="VCTK", multi_speaker=True, max_seq_len=1200,
                             vocoder_speaker="universal")
        if dataset == "LibriTTS":
            return ModelSpec(name="LibriTTS", multi_speaker=True, use_uv=False, max_seq_len=1200,
                             vocoder_speaker="universal")
        raise ValueError(f"unknown dataset preset {dataset!r}")

    @staticmethod
    def from_reference_configs(preprocess_config: dict, model_config: dict, train_config: dict,
                               vocab: int = 361, stats: Optional[dict] = None) -> "ModelSpec":
        """Build from the reference's own config dicts (same keys the reference reads:
        modules.py:110-122, :175-257, :567-572; cmtts.py:24-42; synthesize.py:60-78)."""
        pp = preprocess_config["preprocessing"]
        tr = model_config["transformer"]
        vp = model_config["variance_predictor"]
        ve = model_config["variance_embedding"]
        dn = model_config["denoiser"]
        if pp["pitch"]["pitch_type"] != "cwt":
            raise NotImplementedError("only pitch_type 'cwt' (all shipped configs) is on the hot path")
        if pp["pitch"]["pitch_norm"] != "log":
            raise NotImplementedError("only pitch_norm 'log' (all shipped configs)")
        if pp["energy"]["feature"] != "phoneme_level":
            raise NotImplementedError("only phoneme_level energy (all shipped configs)")
        if ve["energy_quantization"] != "linear":
            raise NotImplementedError("only linear energy quantization (all shipped configs)")
        if tr["ffn_padding"] != "SAME":
            raise NotImplementedError("only ffn_padding SAME (all shipped configs)")
        n_speakers = 0
        if model_config["multi_speaker"] and pp.get("speaker_embedder", "none") == "none":
            with open(os.path.join(preprocess_config["path"]["preprocessed_path"], "speakers.json")) as f:   # cmtts.py:28-34
                n_speakers = len(json.load(f))
        if stats is None:
            p = os.path.join(preprocess_config["path"]["preprocessed_path"], "stats.json")
            with open(p) as f:
                stats = json.load(f)
        e_min, e_max = stats["energy"][:2]
        cm = train_config.get("cm", {})
        return ModelSpec(
            name=str(preprocess_config.get("dataset", "custom")),
            vocab=vocab, hidden=tr["encoder_hidden"], enc_layers=tr["encoder_layer"],
            enc_heads=tr["encoder_head"], ffn_kernel=tr["ffn_kernel_size"], ffn_act=tr["ffn_act"],
            filter_size=vp["filter_size"], dur_layers=vp["dur_predictor_layers"],
            dur_kernel=vp["dur_predictor_kernel"], pred_layers=vp["predictor_layers"],
            pred_kernel=vp["predictor_kernel"], cwt_hidden=vp["cwt_hidden_size"],
            cwt_std_scale=vp["cwt_std_scale"], pitch_bins=ve["pitch_n_bins"],
            energy_bins=ve["energy_n_bins"], use_uv=bool(pp["pitch"]["use_uv"]),
            pitch_norm_eps=float(pp["pitch"]["pitch_norm_eps"]),
            energy_min=float(e_min), energy_max=float(e_max),
            n_mels=pp["mel"]["n_mel_channels"], res_layers=dn["residual_layers"],
            res_channels=dn["residual_channels"], multi_speaker=bool(model_config["multi_speaker"]),
            ext_speaker_dim=((n_speakers + 15) // 16 * 16) if n_speakers else int(model_config.get("external_speaker_dim", 512)),
            n_speakers=n_speakers,
            sigma_min=float(cm.get("sigma_min", 0.002)), sigma_max=float(cm.get("sigma_max", 80.0)),
            sampling_rate=pp["audio"]["sampling_rate"], hop_length=pp["stft"]["hop_length"],
            max_wav_value=float(pp["audio"]["max_wav_value"]), max_seq_len=model_config["max_seq_len"],
            vocoder_speaker=model_config["vocoder"]["speaker"],
        )


# f0 quantiser constants, utils/pitch_tools.py:19-23 (float64 numpy scalars in the reference)
F0_BIN = 256
F0_MAX = 1100.0
F0_MIN = 50.0
F0_MEL_MIN = 1127.0 * math.log(1.0 + F0_MIN / 700.0)
F0_MEL_MAX = 1127.0 * math.log(1.0 + F0_MAX / 700.0)
