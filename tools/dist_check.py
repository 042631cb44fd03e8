"""Multi-GPU check (run under torchrun, NCCL), SURVEY.md §8e / App. D P9:

  * padding="global": the sharded run must reproduce the single-GPU batched run BIT FOR BIT (mel_lens, local mels,
    collated int16 wavs) — small cases with uneven shards (B % world != 0) and, with --full, the C2 bench size
    (LJSpeech, 32 utterances per rank of 80..115 phonemes, T=4: every persistent CTA runs many tiles);
  * padding="local" on length-bucketed shards (what bench.py runs for N > 1): every shard must equal, bit for bit, the
    single-GPU run of the same rows as one batch, and the collation on rank 0 must hold every shard's cropped samples.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [--full]
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmtts_b200 import synthetic  # noqa: E402
from cmtts_b200.config import ModelSpec  # noqa: E402
from cmtts_b200.dist import ShardedSynthesizer, balanced_partition, shard_counts, shard_rows, split_batch  # noqa: E402
from cmtts_b200.synthesize import Pipeline  # noqa: E402


class Replay:
    def __init__(self, tensors, rows):
        self.it, self.rows = iter(tensors), rows

    def randn(self, *shape, device=None, **_):
        return next(self.it)[self.rows].contiguous().to(device)

    def randn_like(self, x):
        return self.randn(*x.shape, device=x.device)


def to_dev(b, dev):
    return (b["texts"].to(dev), b["src_lens"].to(dev), None if b["spker_embeds"] is None else b["spker_embeds"].to(dev))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    full_size = "--full" in sys.argv
    ok = True
    cases = [("VCTK", 6, 8, 30, 4), ("LJSpeech", 4, 20, 45, 2), ("LJSpeech", 2 * world + 1, 10, 30, 1),   # third: uneven shards
             ("VCTK", max(world - 1, 1), 8, 30, 2)]                                                   # fourth: the last rank holds NO rows
    if full_size:
        cases.append(("LJSpeech", 32 * world, 80, 115, 4))
    for ds, B, lo, hi, T in cases:
        spec = ModelSpec.preset(ds)
        sd = synthetic.make_acoustic_state_dict(spec, seed=3)
        ck = synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)
        pipe = Pipeline(spec, sd, ck["generator"], dev)
        batch = synthetic.make_batch(spec, B, lo, hi, seed=11)
        # ---- global padding: bitwise equal to the single-GPU batched run ----
        pre = pipe.model.dpen(batch["texts"], batch["src_lens"], batch["spker_embeds"], None)
        L = pre["cond"].shape[1]
        g = torch.Generator().manual_seed(5)
        noise = [torch.randn(B, 1, L, spec.n_mels, generator=g) for _ in range(T + 1)]   # global noise, row-sliced per rank
        ref = pipe(*to_dev(batch, dev), T=T, generator=Replay(noise, slice(0, B)))
        rows = shard_rows(B, world, rank)
        mine = split_batch(batch, world, rank)
        synth = ShardedSynthesizer(pipe, dist, padding="global", counts=shard_counts(B, world), dst=0)
        out = synth.run(*to_dev(mine, dev), T, generator=Replay(noise, rows), gather=True)
        wav_all, lens_all = synth.collated(out)
        torch.cuda.synchronize()
        same_m = torch.equal(out["mel"], ref["mel"][rows])
        same_w = same_l = True
        if rank == 0:
            same_w = torch.equal(wav_all, ref["wav_i16"])
            same_l = torch.equal(lens_all, ref["mel_lens"])
            print(f"[global] {ds} B={B} (shards {shard_counts(B, world)}) T={T} L={L} world={world}: collated wavs bitwise "
                  f"{same_w}, mel_lens {same_l}, local mels bitwise {same_m}", flush=True)
        else:
            assert wav_all is None and lens_all is None
        ok = ok and same_w and same_l and same_m
        # ---- local padding on length-bucketed shards: each shard == the single-GPU run of its rows ----
        parts = balanced_partition(batch["src_lens"].tolist(), world)
        mine = split_batch(batch, world, rank, rows=parts[rank])
        tmax = int(mine["src_lens"].max()) if len(parts[rank]) else 1
        mine["texts"] = mine["texts"][:, :tmax].contiguous()
        synth = ShardedSynthesizer(pipe, dist, padding="local", counts=[len(p) for p in parts], dst=0)
        torch.manual_seed(1234 + rank)
        out = synth.run(*to_dev(mine, dev), T, gather=True)
        wav_all, lens_all = synth.collated(out)
        torch.manual_seed(1234 + rank)
        alone = pipe(*to_dev(mine, dev), T=T)                      # same rows, same RNG stream, no distributed context
        torch.cuda.synchronize()
        same_s = torch.equal(out["wav_i16"], alone["wav_i16"]) and torch.equal(out["mel"], alone["mel"])
        same_c = True
        if rank == 0:
            n0 = len(parts[0])
            w0 = out["wav_i16"].shape[1]
            same_c = (wav_all.shape[0] == B and torch.equal(wav_all[:n0, :w0], out["wav_i16"])
                      and torch.equal(lens_all[:n0], out["mel_lens"]) and int((lens_all > 0).sum()) == B)
            print(f"[local ] {ds} B={B} (balanced shards {[len(p) for p in parts]}) T={T}: shard == single-GPU run of its rows "
                  f"{same_s}, collation complete {same_c}", flush=True)
        ok = ok and same_s and same_c
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if int(flag.item()) != 0:
        sys.exit(1)
    if rank == 0:
        print("dist_check ok")


if __name__ == "__main__":
    main()
