"""Multi-GPU check (run under torchrun, NCCL): the sharded run with global paddings must reproduce the single-GPU
batched run bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmtts_b200 import synthetic  # noqa: E402
from cmtts_b200.config import ModelSpec  # noqa: E402
from cmtts_b200.dist import ShardedSynthesizer, split_batch  # noqa: E402
from cmtts_b200.synthesize import Pipeline  # noqa: E402


class Replay:
    def __init__(self, tensors, rows):
        self.it, self.rows = iter(tensors), rows

    def randn(self, *shape, device=None, **_):
        return next(self.it)[self.rows].contiguous().to(device)

    def randn_like(self, x):
        return self.randn(*x.shape, device=x.device)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for ds, B, lo, hi, T in [("VCTK", 6, 8, 30, 4), ("LJSpeech", 4, 20, 45, 2)]:
        spec = ModelSpec.preset(ds)
        sd = synthetic.make_acoustic_state_dict(spec, seed=3)
        ck = synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)
        pipe = Pipeline(spec, sd, ck["generator"], dev)
        batch = synthetic.make_batch(spec, B, lo, hi, seed=11)
        # noise for the GLOBAL batch, row-sliced per rank (L is only known after the pre-pass: draw generously)
        full = Pipeline(spec, sd, ck["generator"], dev)
        pre = full.model.dpen(batch["texts"], batch["src_lens"], batch["spker_embeds"], None)
        L = pre["cond"].shape[1]
        g = torch.Generator().manual_seed(5)
        noise = [torch.randn(B, 1, L, spec.n_mels, generator=g) for _ in range(T + 1)]
        ref = full(batch["texts"], batch["src_lens"], batch["spker_embeds"], T=T, generator=Replay(noise, slice(0, B)))
        mine = split_batch(batch, world, rank)
        from cmtts_b200.dist import shard_rows
        rows = shard_rows(B, world, rank)
        synth = ShardedSynthesizer(pipe, dist)
        out = synth.run(mine["texts"].to(dev), mine["src_lens"].to(dev),
                        None if mine["spker_embeds"] is None else mine["spker_embeds"].to(dev), T,
                        generator=Replay(noise, rows), gather=True)
        torch.cuda.synchronize()
        same_w = torch.equal(out["wav_i16_all"], ref["wav_i16"])
        same_l = torch.equal(out["mel_lens_all"], ref["mel_lens"])
        same_m = torch.equal(out["mel"], ref["mel"][rows])
        if rank == 0:
            print(f"{ds} B={B} T={T} world={world}: wavs bitwise {same_w}, mel_lens {same_l}, local mels bitwise {same_m}", flush=True)
        ok = ok and same_w and same_l and same_m
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if int(flag.item()) != 0:
        sys.exit(1)
    if rank == 0:
        print("dist_check ok")


if __name__ == "__main__":
    main()
