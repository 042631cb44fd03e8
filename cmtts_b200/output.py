"""Output stage of the synthesis path: mels -> int16 wavs -> files.

Mirrors `synth_samples` (reference utils/tools.py:566-607) and its file naming; what changes is the
schedule, not the results:
  * `mel_lens` is read back ONCE (the reference calls `.item()` twice per utterance, tools.py:577-578);
  * the x32768 scaling, truncating int16 cast (utils/model.py:195-198) run on the device, so only int16
    samples cross PCIe, into a pinned staging buffer;
  * WAV files are written by a small thread pool while the next batch is being synthesized
    (`AsyncWavWriter`); the bytes are `scipy.io.wavfile.write`'s, as in the reference;
  * the per-utterance matplotlib PNG (tools.py:582-592), which dominates the reference's wall clock once
    the model is fast, is opt-in (`plot=True`, needs matplotlib).
"""
from __future__ import annotations

import os
from concurrent.futures import Future, ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np
import torch

from .vocoder import Generator


def output_name(basename: str, args, multi_speaker: bool, ext: str) -> str:
    """`{basename}_{speaker_id}{tag}.ext` for multi-speaker single-sentence runs, `{basename}{tag}.ext` otherwise
    (utils/tools.py:583-585, :603-605)."""
    tag = "_teacher_forced" if getattr(args, "teacher_forced", False) else ""
    if multi_speaker and getattr(args, "mode", None) == "single":
        return "{}_{}{}.{}".format(basename, getattr(args, "speaker_id", None), tag, ext)
    return "{}{}.{}".format(basename, tag, ext)


class AsyncWavWriter:
    """Writes 16-bit PCM WAV files on worker threads; `close()` (or leaving the `with` block) waits for all of
    them and re-raises the first error."""

    def __init__(self, workers: int = 4):
        self._pool = ThreadPoolExecutor(max_workers=max(1, int(workers)))
        self._pending: List[Future] = []

    @staticmethod
    def _write(path: str, rate: int, wav: np.ndarray) -> str:
        from scipy.io import wavfile
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        wavfile.write(path, rate, wav)
        return path

    def submit(self, path: str, rate: int, wav: np.ndarray) -> Future:
        if wav.dtype != np.int16:
            raise TypeError("AsyncWavWriter writes int16 PCM")
        f = self._pool.submit(self._write, path, int(rate), np.ascontiguousarray(wav).copy())
        self._pending.append(f)
        return f

    def flush(self) -> List[str]:
        done = [f.result() for f in self._pending]
        self._pending = []
        return done

    def close(self) -> List[str]:
        try:
            return self.flush()
        finally:
            self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class _Pinned:
    """Grow-only pinned int16 staging buffer for the device -> host copy of a batch of wavs."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def get(self, shape) -> torch.Tensor:
        n = int(np.prod(shape))
        if self.buf is None or self.buf.numel() < n:
            self.buf = torch.empty(n, dtype=torch.int16, pin_memory=torch.cuda.is_available())
        return self.buf[:n].view(*shape)


_PINNED = _Pinned()


def wavs_from_mels(mel_blc: torch.Tensor, mel_lens: Sequence[int], vocoder: Generator, hop_length: int,
                   max_wav_value: float = 32768.0) -> List[np.ndarray]:
    """(B, L, 80) channels-last mels (the sampler's layout) -> list of int16 arrays cropped to mel_len * hop
    (utils/model.py:187-205; the whole padded batch goes through the vocoder, exactly as the reference does)."""
    _, w16 = vocoder.run(mel_blc, want_float=False, want_int16=True, max_wav_value=max_wav_value)
    host = _PINNED.get(tuple(w16.shape))
    host.copy_(w16, non_blocking=True)
    torch.cuda.current_stream(w16.device).synchronize()
    arr = host.numpy()
    return [arr[i, : int(n) * hop_length].copy() for i, n in enumerate(mel_lens)]


def synth_samples(args, targets, predictions, vocoder: Generator, model_config, preprocess_config, path: str,
                  diffusion=None, plot: bool = False, writer: Optional[AsyncWavWriter] = None) -> List[str]:
    """utils/tools.py:566-607: `predictions` is `CMTotalTTSSynthesize.synthesize`'s 12-slot list
    ([0] mels (B, L, 80), [10] src_lens, [11] mel_lens), `targets[0]` the basenames.  Returns the WAV paths.
    With `writer` the files are written asynchronously (call `writer.flush()` / `close()` before reading them)."""
    if getattr(args, "model", "naive") == "aux":
        raise NotImplementedError("the 'aux' model variant is not on the inference path (tools.py:572-574)")
    multi_speaker = bool(model_config["multi_speaker"])
    basenames = list(targets[0])
    mels = predictions[0]
    mel_lens = torch.as_tensor(predictions[11]).cpu().tolist()          # one read-back
    out_dir = os.path.join(path, str(getattr(args, "restore_step", 0)))
    os.makedirs(out_dir, exist_ok=True)
    if plot:
        _plot_mels(mels, mel_lens, basenames, args, multi_speaker, out_dir)
    pre = preprocess_config["preprocessing"]
    wavs = wavs_from_mels(mels, mel_lens, vocoder, int(pre["stft"]["hop_length"]), float(pre["audio"]["max_wav_value"]))
    rate = int(pre["audio"]["sampling_rate"])
    own = writer is None
    w = writer or AsyncWavWriter(workers=1)
    paths = []
    for wav, basename in zip(wavs, basenames):
        p = os.path.join(out_dir, output_name(basename, args, multi_speaker, "wav"))
        w.submit(p, rate, wav)
        paths.append(p)
    if own:
        w.close()
    return paths


def _plot_mels(mels, mel_lens, basenames, args, multi_speaker, out_dir):   # pragma: no cover - needs matplotlib
    try:
        import matplotlib
        matplotlib.use("Agg")
        from matplotlib import pyplot as plt
    except ImportError as e:
        raise RuntimeError("plot=True needs matplotlib (reference utils/tools.py:10-11)") from e
    host = mels.detach().float().cpu().numpy()
    for i, basename in enumerate(basenames):
        mel = host[i, : int(mel_lens[i])].T
        fig, ax = plt.subplots(1, 1, figsize=(8, 4), squeeze=True)
        ax.imshow(mel, origin="lower")
        ax.set_aspect(2.5, adjustable="box")
        ax.set_ylim(0, mel.shape[0])
        ax.set_title("Synthetized Spectrogram", fontsize="medium")
        ax.tick_params(labelsize="x-small", left=False, labelleft=False)
        ax.set_anchor("W")
        plt.savefig(os.path.join(out_dir, output_name(basename, args, multi_speaker, "png")))
        plt.close(fig)
