"""Synthesis driver with the reference's protocol (synthesize.py:35-153, p_rtf_cm.py:174-230).

    tool = CMTotalTTSSynthesize(model_path, restore_step, args, preprocess_config, model_config, train_config)
    out_put = tool.synthesize(batch)        # batch = the reference's 7-tuple; out_put[0] = mel (B, L, 80),
                                            # out_put[10] = src_lens, out_put[11] = mel_lens

Differences from the reference that do not change results: the checkpoint is loaded once (the
reference reloads it for every batch, synthesize.py:203), the conditioner is computed once per
batch instead of T+1 times, and the pre-pass result is reused by the sampler.
`Pipeline` is the whole hot path (host batch -> int16 wavs) used by bench.py and the multi-GPU
driver.
"""
from __future__ import annotations

import argparse
import os
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .config import HifiGanSpec, ModelSpec
from .model import CMTotalTTS, KarrasDenoiser, create_model_and_diffusion_tts
from .sampler import karras_sample_tts, sampler_plan
from .vocoder import Generator, vocoder_infer


def checkpoint_path(model_path: str, step: int, kind: str = "model") -> str:
    """`<model_path>/CMDenoiserTTS/{model|target_model|teacher_model}{step:06d}.pt` or `ema_<rate>_{step:06d}.pt`
    (written by train_util.py:890-917; synthesize.py:44-48 reads the first)."""
    if kind in ("model", "target_model", "teacher_model"):
        name = "{}{:06d}.pt".format(kind, int(step))
    elif kind.startswith("ema_"):
        name = "{}_{:06d}.pt".format(kind, int(step))
    else:
        raise ValueError(f"unknown checkpoint kind {kind!r}")
    return os.path.join(model_path, "CMDenoiserTTS", name)


def to_device(data, device):
    """utils/tools.py:103-112 — the 7-tuple inference batch:
    (ids, raw_texts, speakers, texts, src_lens, max_src_len, spker_embeds)."""
    if len(data) != 7:
        raise ValueError("inference batches are 7-tuples (utils/tools.py:103-112)")
    ids, raw_texts, speakers, texts, src_lens, max_src_len, spker_embeds = data
    speakers = torch.as_tensor(np.asarray(speakers)).long().to(device)
    texts = torch.as_tensor(np.asarray(texts)).long().to(device)
    src_lens = torch.as_tensor(np.asarray(src_lens)).to(device)
    if spker_embeds is not None:
        spker_embeds = torch.as_tensor(np.asarray(spker_embeds)).float().to(device)
    return [ids, raw_texts, speakers, texts, src_lens, max_src_len, spker_embeds]


class CMTotalTTSSynthesize:
    """synthesize.py:35-153."""

    def __init__(self, model_path, model_step_num, args, preprocess_config, model_config, train_config,
                 p_control=1.0, e_control=1.0, d_control=1.0, device=None, spec: Optional[ModelSpec] = None,
                 checkpoint: str = "model", forward_controls: bool = False):
        """`checkpoint` picks which of the trainer's files is loaded (train_util.py:890-917): "model" (what the
        reference's synthesize.py:44-48 reads), "target_model", "teacher_model" or "ema_<rate>" — all are flat
        CMTotalTTS state_dicts.  `forward_controls=True` hands p/e/d_control to the variance adaptor; the reference
        parses them (synthesize.py:275-292) but never forwards them (:96-102), so the default keeps them inert."""
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.CMDenoiserTTS_path = checkpoint_path(model_path, model_step_num, checkpoint)
        self.forward_controls = bool(forward_controls)
        self.args = args
        self.train_config = train_config
        self.p_control, self.e_control, self.d_control = p_control, e_control, d_control
        self.model, self.diffusion = self.load_cm_model(args, preprocess_config, model_config, train_config, spec)
        self.duration_pitch_energy_net, self.denoise_net = self.model.get_segmentation_model()

    def load_cm_model(self, args, preprocess_config, model_config, train_config, spec=None):
        cm = dict(train_config.get("cm", {})) if train_config else {}
        mode = cm.get("training_mode", "consistency_distillation")
        if mode == "progdist":
            distillation = False
        elif "consistency" in mode:
            distillation = True
        else:
            raise ValueError(f"unknown training mode {mode}")   # synthesize.py:66
        if spec is not None:
            model = CMTotalTTS(spec=spec)
            diffusion = KarrasDenoiser(sigma_data=0.5, sigma_max=spec.sigma_max, sigma_min=spec.sigma_min,
                                       distillation=distillation)
        else:
            model, diffusion = create_model_and_diffusion_tts(
                use_fp16=cm.get("use_fp16", False), weight_schedule=cm.get("weight_schedule", "uniform"),
                tts_model_config={"args": args, "train_config": train_config,
                                  "preprocess_config": preprocess_config, "model_config": model_config},
                sigma_min=cm.get("sigma_min", 0.002), sigma_max=cm.get("sigma_max", 80.0),
                distillation=distillation, loss_norm=cm.get("loss_norm", "mel_loss"))
        model.load_state_dict(torch.load(self.CMDenoiserTTS_path, map_location="cpu", weights_only=True))
        model.to(self.device)
        model.eval()
        return model, diffusion

    def synthesize(self, batch, T: Optional[int] = None, generator=None, trace=None, ref_audio: Optional[str] = None,
                   speaker_ckpt: Optional[str] = None):
        """`ref_audio`: the zero-shot variant (synthesize_zeroshot_lj.py:89-102 / synthesize_zeroshot_vctk.py): the speaker
        embedding of every utterance of the batch is the DeepSpeaker embedding of this recording (computed on the GPU by
        cmtts_b200.speaker_encoder; `speaker_ckpt` = the Keras checkpoint, default the reference's relative path)
        instead of the batch's precomputed one."""
        T = int(T if T is not None else getattr(self.args, "T", 1))
        kw = {"speakers": batch[2], "texts": batch[3], "src_lens": batch[4], "spker_embeds": batch[-1]}
        if ref_audio is not None:
            from .speaker_encoder import get_deep_speaker_emb
            kw["spker_embeds"] = get_deep_speaker_emb(filepath=ref_audio, batch_size=batch[2].size(0), device=batch[2].device,
                                                      ckpt_path=speaker_ckpt)
        ctl = {"p_control": self.p_control, "e_control": self.e_control, "d_control": self.d_control} \
            if self.forward_controls else {}
        out_dict = self.duration_pitch_energy_net(**kw, **ctl)
        batch_size, seq_len, _ = out_dict["cond"].size()
        sampler, steps, ts = sampler_plan(T)
        s = self.model.spec
        sample = karras_sample_tts(
            diffusion=self.diffusion, model=self.model, shape=(batch_size, 1, seq_len, s.n_mels),
            model_kwargs=kw, device=self.device, sigma_max=self.diffusion.sigma_max,
            sigma_min=self.diffusion.sigma_min, sampler=sampler, steps=steps, ts=ts,
            generator=generator, cond_dict=out_dict, trace=trace)
        out_put = [None] * 12
        out_put[0] = sample
        out_put[10] = kw["src_lens"]
        out_put[11] = out_dict["mel_lens"]
        self.last_out_dict = out_dict
        return out_put


class _GraphSlot:
    """One captured piece of the step: eager on its first call (warm-up: sizes the grow-only workspaces, sets kernel
    attributes), captured into a CUDA graph on the second, replayed from then on."""

    _gen = [0]

    def __init__(self):
        self.calls = 0
        self.graph = None
        self.static_in = None
        self.out = None
        self.failed = False
        self.n_kernels = 0
        _GraphSlot._gen[0] += 1
        self.gen = _GraphSlot._gen[0]          # names this slot's static buffers (a tail graph reads its head's by address)


class Pipeline:
    """Whole hot path on one GPU: host phoneme ids (+ speaker embeddings) -> mels -> int16 wavs.

    Mirrors what p_rtf_cm.py:174-226 strings together (encoder + variance adaptor, T solver steps,
    HiFi-GAN on the whole padded batch, x32768 -> int16, crop to mel_len * hop).

    graphs=True: the step is replayed from two CUDA graphs per shape — the token-rate part before the single host
    read of the output length, keyed on (B, Tsrc, T), and everything after it (length regulator ... int16 wavs, ~130
    launches at T = 1), keyed on (B, Tsrc, L, T) — instead of being enqueued launch by launch.  It pays when the step is
    launch-latency bound (single utterances, small batches: SURVEY.md 7.6); a big batch is GPU-bound either way.  Shapes are
    not bucketed: the reference's results depend on the exact padded length (SURVEY.md 7 "padding semantics"), so a new
    (B, Tsrc, L) runs eagerly once, is captured on its second appearance, and the `graph_cache` most recent shapes are
    kept.  Only the built-in device RNG can be captured: a call with a `generator` (noise replay) or a trace runs eagerly.
    Results are fresh tensors either way (the graphs' static outputs are cloned)."""

    def __init__(self, spec: ModelSpec, acoustic_sd: Dict[str, torch.Tensor], hifigan_sd: Dict[str, torch.Tensor],
                 device, distillation: bool = True, precision: str = "tc", tc_frontend: bool = True,
                 graphs: bool = False, graph_cache: int = 16):
        self.spec = spec
        self.precision = precision
        self.device = torch.device(device)
        self.model = CMTotalTTS(spec=spec, precision=precision).load_state_dict(acoustic_sd).to(self.device)
        self.model.tc_frontend = bool(tc_frontend)
        self.diffusion = KarrasDenoiser(sigma_data=spec.sigma_data, sigma_max=spec.sigma_max,
                                        sigma_min=spec.sigma_min, rho=spec.rho, distillation=distillation)
        self.vocoder = Generator(hspec=spec.hifigan, precision=precision).load_state_dict(hifigan_sd).to(self.device)
        self.graphs = bool(graphs)
        self.graph_cache = int(graph_cache)
        self._head_slots: "Dict[tuple, _GraphSlot]" = {}
        self._tail_slots: "Dict[tuple, _GraphSlot]" = {}
        self._t_cache: Dict[tuple, torch.Tensor] = {}
        self.graph_replays = 0
        self.graph_kernel_launches = 0     # kernels of this library executed through graph replays (not counted by the C side)

    def acoustic(self, texts, src_lens, spker_embeds, T: int, generator=None, l_max_hook=None, trace=None):
        out = self.model.dpen(texts, src_lens, spker_embeds, None, l_max_hook=l_max_hook)
        B, L, _ = out["cond"].shape
        sampler, steps, ts = sampler_plan(T)
        kw = {"texts": texts, "src_lens": src_lens, "spker_embeds": spker_embeds}
        mel = karras_sample_tts(self.diffusion, self.model, (B, 1, L, self.spec.n_mels), steps=steps,
                                model_kwargs=kw, device=self.device, sigma_min=self.spec.sigma_min,
                                sigma_max=self.spec.sigma_max, sampler=sampler, ts=ts, generator=generator,
                                cond_dict=out, trace=trace)
        return mel, out

    def __call__(self, texts, src_lens, spker_embeds=None, T: int = 1, generator=None, want_float_wav: bool = False,
                 l_max_hook=None):
        if self.graphs and generator is None and not want_float_wav and texts.shape[0] > 0:
            return self._call_graphed(texts, src_lens, spker_embeds, T, l_max_hook)
        mel, out = self.acoustic(texts, src_lens, spker_embeds, T, generator, l_max_hook)
        wav, w16 = self.vocoder.run(mel, want_float=want_float_wav, want_int16=True,
                                    max_wav_value=self.spec.max_wav_value)
        return {"mel": mel, "mel_lens": out["mel_lens"], "wav_i16": w16, "wav": wav, "dpen": out}

    # ---- CUDA-graph path ---------------------------------------------------------------------------------------------
    def _timesteps(self, B: int, T: int) -> torch.Tensor:
        """Device copy of the reference's `rescaled_t` for this T's evaluation sigma (uploaded once per (B, T))."""
        from .sampler import evaluation_sigma, rescaled_timesteps
        t = self._t_cache.get((B, T))
        if t is None:
            sampler, steps, _ = sampler_plan(T)
            sigma = evaluation_sigma(sampler, steps, self.spec.sigma_min, self.spec.sigma_max, self.diffusion.rho)
            t = self._t_cache[(B, T)] = rescaled_timesteps(B, sigma).to(self.device)
        return t

    def _head(self, texts, src_lens, spk, T):
        head = self.model.dpen_head(texts, src_lens, spk)
        steps = self.model.prepare_steps(self._timesteps(texts.shape[0], T), head["spk"])
        return head, steps

    def _tail(self, head, steps, L, local_max, extra, T):
        out = self.model.dpen_tail(head, L, local_max, 1.0, extra)
        B = head["texts"].shape[0]
        sampler, nsteps, ts = sampler_plan(T)
        kw = {"texts": head["texts"], "src_lens": head["src_lens"], "spker_embeds": None}
        mel = karras_sample_tts(self.diffusion, self.model, (B, 1, L, self.spec.n_mels), steps=nsteps, model_kwargs=kw,
                                device=self.device, sigma_min=self.spec.sigma_min, sigma_max=self.spec.sigma_max,
                                sampler=sampler, ts=ts, cond_dict=out, prepared_steps=steps)
        _, w16 = self.vocoder.run(mel, want_float=False, want_int16=True, max_wav_value=self.spec.max_wav_value)
        return out, mel, w16

    def _slot(self, table, key):
        slot = table.pop(key, None) or _GraphSlot()
        table[key] = slot                                  # most recently used last
        while len(table) > self.graph_cache:
            table.pop(next(iter(table)))
        return slot

    def _run_slot(self, slot, fn, inputs):
        """inputs: tuple of tensors (or None) handed to fn; copied into the slot's static buffers when a graph exists."""
        slot.calls += 1
        if slot.graph is not None:
            for dst, src in zip(slot.static_in, inputs):
                if dst is not None and dst is not src:
                    dst.copy_(src, non_blocking=True)
            slot.graph.replay()
            self.graph_replays += 1
            self.graph_kernel_launches += slot.n_kernels
            return slot.out
        if slot.calls < 2 or slot.failed:
            return fn(*inputs)                             # first sight of this shape: eager (also the warm-up)
        static_in = tuple(None if t is None else t.clone() for t in inputs)
        torch.cuda.current_stream(self.device).synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = self.model.lib.cmtts_launch_count()
        try:
            with torch.cuda.graph(g):
                out = fn(*static_in)
        except Exception as e:                             # not capturable on this driver: say so once, stay eager
            slot.failed = True
            import warnings
            warnings.warn(f"cmtts_b200: CUDA-graph capture failed ({type(e).__name__}: {e}); running this shape eagerly")
            torch.cuda.synchronize(self.device)
            return fn(*inputs)
        slot.static_in, slot.out, slot.graph = static_in, out, g
        slot.n_kernels = int(self.model.lib.cmtts_launch_count() - n0)     # kernel nodes of this library in the graph
        g.replay()                                         # capture only records: run it once for this call's result
        self.graph_replays += 1
        self.graph_kernel_launches += slot.n_kernels
        return slot.out

    def _call_graphed(self, texts, src_lens, spker_embeds, T, l_max_hook):
        m = self.model
        m._ready()
        with torch.cuda.device(self.device):
            texts, src_lens, spk = m.prepare_inputs(texts, src_lens, spker_embeds)
            B, Tsrc = texts.shape
            hslot = self._slot(self._head_slots, (B, Tsrc, T))
            head, steps = self._run_slot(hslot, lambda a, b, c: self._head(a, b, c, T), (texts, src_lens, spk))
            local_max, extra = m.read_lengths(head, l_max_hook)          # the one host round trip
            L = local_max
            tslot = self._slot(self._tail_slots, (B, Tsrc, L, T, hslot.gen))
            if hslot.graph is None:
                # the tail graph would read the head's tensors by address: only valid once the head is static too
                out, mel, w16 = self._tail(head, steps, L, local_max, extra, T)
            else:
                out, mel, w16 = self._run_slot(tslot, lambda: self._tail(head, steps, L, local_max, extra, T), ())
            # results are fresh tensors: the graphs' static outputs are overwritten by the next call of the same shape
            # (the `dpen` dict is handed out as is: in graph mode its entries alias those static buffers)
            fresh = tslot.graph is not None
            return {"mel": mel.clone() if fresh else mel,
                    "mel_lens": out["mel_lens"].clone() if hslot.graph is not None else out["mel_lens"],
                    "wav_i16": w16.clone() if fresh else w16, "wav": None, "dpen": out}

    def crop(self, w16_host: np.ndarray, mel_lens: Sequence[int]) -> List[np.ndarray]:
        """utils/model.py:201-203."""
        return [w16_host[i, : int(n) * self.spec.hop_length] for i, n in enumerate(mel_lens)]


def rtf_like_reference(pipe: Pipeline, texts, src_lens, spker_embeds, T: int, out_wav_path: Optional[str] = None):
    """RTF exactly as p_rtf_cm.py:190-230 defines it: the timer starts AFTER the encoder/variance
    pre-pass, and stops after sampling, vocoding the whole batch, D2H, int16 conversion and writing
    the FIRST wav; divided by the duration of that first utterance.  Returns (rtf_ref, rtf_total, elapsed)."""
    dev = pipe.device
    out = pipe.model.dpen(texts, src_lens, spker_embeds, None)
    torch.cuda.synchronize(dev)
    t0 = time.time()
    B, L, _ = out["cond"].shape
    sampler, steps, ts = sampler_plan(T)
    kw = {"texts": texts, "src_lens": src_lens, "spker_embeds": spker_embeds}
    mel = karras_sample_tts(pipe.diffusion, pipe.model, (B, 1, L, pipe.spec.n_mels), steps=steps, model_kwargs=kw,
                            device=dev, sigma_min=pipe.spec.sigma_min, sigma_max=pipe.spec.sigma_max,
                            sampler=sampler, ts=ts, cond_dict=out)
    _, w16 = pipe.vocoder.run(mel, want_float=False, want_int16=True, max_wav_value=pipe.spec.max_wav_value)
    host = w16.cpu().numpy()
    mel_lens = out["mel_lens"].cpu().tolist()
    wavs = pipe.crop(host, mel_lens)
    if out_wav_path is not None:
        from scipy.io import wavfile
        wavfile.write(out_wav_path, pipe.spec.sampling_rate, wavs[0])
    elapsed = time.time() - t0
    dur0 = mel_lens[0] * pipe.spec.hop_length / pipe.spec.sampling_rate
    total = sum(mel_lens) * pipe.spec.hop_length / pipe.spec.sampling_rate
    return elapsed / dur0, elapsed / total, elapsed


# ------------------------------------------------------------------------------------------------
# command line: the reference's `python synthesize.py ...` (synthesize.py:227-397), same flags
# ------------------------------------------------------------------------------------------------
def get_configs_of(dataset: str, config_dir: str = "./config"):
    """utils/tools.py:25-34: `<config_dir>/<dataset>/{preprocess,model,train}.yaml` (cwd-relative in the reference)."""
    import yaml

    out = []
    for name in ("preprocess", "model", "train"):
        with open(os.path.join(config_dir, dataset, name + ".yaml")) as f:
            out.append(yaml.load(f, Loader=yaml.FullLoader))
    return tuple(out)


def build_arg_parser() -> argparse.ArgumentParser:
    """The reference's flags (synthesize.py:229-310) with their names and defaults.  Its copy-pasted help strings and
    the float defaults of `--model_path` / `--T` are not reproduced; `--model_path` is required here (the reference
    crashes in os.path.join without it).  Added: --config_dir, --hifigan_dir, --device, --checkpoint,
    --forward_controls, --batch_size (the reference hard-codes 8, :368)."""
    p = argparse.ArgumentParser(prog="python -m cmtts_b200.synthesize")
    p.add_argument("--restore_step", type=int, required=True)
    p.add_argument("--path_tag", type=str, default="")
    p.add_argument("--model", type=str, choices=["naive", "aux", "shallow"], default="naive", help="training model type")
    p.add_argument("--teacher_forced", action="store_true")
    p.add_argument("--mode", type=str, choices=["batch", "single"], required=True,
                   help="synthesize a source file (batch) or one sentence (single)")
    p.add_argument("--source", type=str, default=None, help="file in the format of train.txt / val.txt (batch mode)")
    p.add_argument("--text", type=str, default=None, help="raw text (single mode)")
    p.add_argument("--speaker_id", type=str, default="p225", help="speaker for multi-speaker models (single mode)")
    p.add_argument("--dataset", type=str, required=True, help="config/<dataset>/*.yaml")
    p.add_argument("--pitch_control", type=float, default=1.0)
    p.add_argument("--energy_control", type=float, default=1.0)
    p.add_argument("--duration_control", type=float, default=1.0)
    p.add_argument("--result_path", type=str, default=None, help="output directory (default: train.yaml path.result_path)")
    p.add_argument("--model_path", type=str, required=True, help="directory that holds CMDenoiserTTS/model<step>.pt")
    p.add_argument("--T", type=int, default=1, choices=[1, 2, 4], help="consistency sampling steps")
    p.add_argument("--config_dir", type=str, default="./config")
    p.add_argument("--hifigan_dir", type=str, default="hifigan", help="directory of config.json + generator_*.pth.tar")
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--checkpoint", type=str, default="model",
                   help="model | target_model | teacher_model | ema_<rate> (train_util.py:890-917)")
    p.add_argument("--forward_controls", action="store_true",
                   help="hand the p/e/d controls to the variance adaptor (the reference parses but never forwards them)")
    p.add_argument("--batch_size", type=int, default=8)
    return p


def check_args(args) -> None:
    """synthesize.py:312-320 (asserts in the reference)."""
    if args.mode == "batch":
        if args.text is not None:
            raise ValueError("--text is for --mode single")
        if args.teacher_forced:
            raise NotImplementedError("--teacher_forced needs ground-truth features (training data path, out of scope)")
        if args.source is None:
            raise ValueError("--mode batch needs --source")
    if args.mode == "single":
        if args.source is not None or args.text is None or args.teacher_forced:
            raise ValueError("--mode single needs --text and neither --source nor --teacher_forced")


def prepare_batches(args, preprocess_config, model_config, g2p=None):
    """synthesize.py:358-394: the list of 7-tuple batches for either mode."""
    from .frontend import TextDataset, single_batch

    if args.mode == "batch":
        return list(TextDataset(args.source, preprocess_config, model_config).batches(args.batch_size))
    return [single_batch(args.text, args.speaker_id, preprocess_config, model_config, g2p)]


def main(argv=None) -> List[str]:
    """Text in, WAV files out, through the CUDA hot path; returns the paths written.  Unlike the reference the
    checkpoint and vocoder are loaded once, not per batch (synthesize.py:203)."""
    from .output import AsyncWavWriter, synth_samples
    from .vocoder import get_vocoder

    args = build_arg_parser().parse_args(argv)
    check_args(args)
    preprocess_config, model_config, train_config = get_configs_of(args.dataset, args.config_dir)
    path_tag = "_{}".format(args.path_tag) if args.path_tag != "" else ""
    result_path = args.result_path or (train_config["path"]["result_path"] + "_{}{}".format(args.model, path_tag))
    # only len() of cwt_scales is read on this path (pitch_tools.py:246); the reference computes it with pycwt (:333-336)
    preprocess_config["preprocessing"]["pitch"].setdefault("cwt_scales", [0.0] * 10)
    batchs = prepare_batches(args, preprocess_config, model_config)
    if not torch.cuda.is_available():
        raise _lib.CmttsError("cmtts_b200.synthesize needs a CUDA device (there is no CPU fallback for the hot path)")
    device = torch.device(args.device)
    speaker = model_config["vocoder"]["speaker"]
    vocoder = get_vocoder(model_config, device,
                          checkpoint_path=os.path.join(args.hifigan_dir, f"generator_{speaker}.pth.tar"),
                          hifigan_config=os.path.join(args.hifigan_dir, "config.json"))
    tool = CMTotalTTSSynthesize(args.model_path, args.restore_step, args, preprocess_config, model_config, train_config,
                                p_control=args.pitch_control, e_control=args.energy_control,
                                d_control=args.duration_control, device=device, checkpoint=args.checkpoint,
                                forward_controls=args.forward_controls)
    written: List[str] = []
    with torch.no_grad(), AsyncWavWriter() as writer:
        for batch in batchs:
            batch = to_device(batch, device)
            out_put = tool.synthesize(batch)
            written += synth_samples(args, batch, out_put, vocoder, model_config, preprocess_config, result_path,
                                     tool.diffusion, writer=writer)
    return written


if __name__ == "__main__":
    for path in main():
        print(path)
