"""Host-side logic of the utterance sharding (cmtts_b200/dist.py) with world_size 2 over gloo on CPU.
The per-rank compute is stood in by the oracle so that the test runs without a GPU: what is under
test is the splitting, the global-L_max exchange and the gather, i.e. that the sharded run with
global paddings reproduces the single-batch result bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cmtts_b200 import synthetic
from cmtts_b200.config import ModelSpec
from cmtts_b200.dist import (GlobalMax, LocalMaxWithWire, balanced_partition, gather_rows, shard_counts, shard_rows,
                             split_batch)


def test_shard_rows_is_a_balanced_partition():
    for n in (1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                s = shard_rows(n, w, r)
                seen += list(range(n))[s]
            assert seen == list(range(n))
            sizes = [len(range(n)[shard_rows(n, w, r)]) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_split_batch_keeps_global_token_padding():
    spec = ModelSpec.preset("VCTK")
    b = synthetic.make_batch(spec, 5, 4, 20, seed=3)
    parts = [split_batch(b, 2, r) for r in range(2)]
    assert parts[0]["texts"].shape[1] == parts[1]["texts"].shape[1] == b["texts"].shape[1]
    assert torch.equal(torch.cat([p["texts"] for p in parts]), b["texts"])
    assert torch.equal(torch.cat([p["spker_embeds"] for p in parts]), b["spker_embeds"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    from oracle import cmtts_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    spec = ModelSpec.preset("LJSpeech")
    sd = synthetic.make_acoustic_state_dict(spec, seed=5)
    W = O.Weights(sd)
    full = synthetic.make_batch(spec, 4, 6, 16, seed=11)
    mine = split_batch(full, world, rank)
    hook = GlobalMax(dist, "cpu")
    with torch.no_grad():
        # local pre-pass gives the local L_max; the hook turns it into the global one
        local = O.dpen(W, spec, **mine)
        L = hook(int(local["mel_lens"].max()))
        out = O.dpen(W, spec, max_mel_len=L, **mine)
    cond_all = gather_rows(dist, out["cond"])
    lens_all = gather_rows(dist, out["mel_lens"])
    if rank == 0:
        with torch.no_grad():
            ref = O.dpen(W, spec, **full)
        q.put((bool(torch.equal(cond_all, ref["cond"])), bool(torch.equal(lens_all, ref["mel_lens"])), L,
               int(ref["cond"].shape[1]), hook.calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_batch_bitwise():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    cond_eq, lens_eq, L, L_ref, calls = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert L == L_ref           # global L_max exchanged
    assert lens_eq and cond_eq  # bitwise identical to the single-batch run
    assert calls == 1


def test_balanced_partition_is_a_partition_and_cuts_padding():
    import random
    rnd = random.Random(0)
    for n, w in ((1, 2), (5, 8), (32, 1), (64, 2), (256, 8), (33, 4)):
        lens = [rnd.randint(80, 115) for _ in range(n)]
        parts = balanced_partition(lens, w)
        assert len(parts) == w and sorted(i for p in parts for i in p) == list(range(n))
        cost = max((len(p) * max(lens[i] for i in p)) if p else 0 for p in parts)
        contiguous = max(len(range(n)[shard_rows(n, w, r)]) * max(lens) for r in range(w))   # global-padding shards
        assert cost <= contiguous
        # shards are length-sorted: no shard's shortest row is longer than a previous shard's longest
        tops = [max(lens[i] for i in p) for p in parts if p]
        assert tops == sorted(tops, reverse=True)
    # 8 shards of 256 utterances: padded/valid drops well below the single-batch figure
    lens = [rnd.randint(80, 115) for _ in range(256)]
    parts = balanced_partition(lens, 8)
    padded = sum(len(p) * max(lens[i] for i in p) for p in parts)
    assert padded / sum(lens) < 1.04 < 256 * max(lens) / sum(lens)


def _worker_uneven(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 5                                              # 5 rows over 2 ranks: 3 + 2; and 1 row over 2 ranks: 1 + 0
    full = torch.arange(n * 6, dtype=torch.int16).view(n, 6)
    counts = shard_counts(n, world)
    mine = full[shard_rows(n, world, rank)]
    everywhere = gather_rows(dist, mine, counts)                       # all_gather, uneven shards
    at0 = gather_rows(dist, mine, counts, dst=0)                       # gather to one rank
    pend = gather_rows(dist, mine.float(), counts, dst=0, async_op=True)
    late = pend.result()
    one = torch.ones(1, 3)
    c1 = shard_counts(1, world)
    solo = gather_rows(dist, one[shard_rows(1, world, rank)], c1, dst=0)   # a rank with ZERO rows takes part
    hook = LocalMaxWithWire(dist, "cpu")
    both = hook(torch.tensor(10 + rank))
    ok = bool(torch.equal(everywhere, full)) and both.tolist() == [10 + rank, 10 + world - 1]
    if rank == 0:
        ok = ok and torch.equal(at0, full) and torch.equal(late, full.float()) and torch.equal(solo, one)
        q.put(ok)
    else:
        assert at0 is None and late is None and solo is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_uneven_shards_gather_to_one_rank_gloo():
    """B % world != 0 and B < world (all_gather of mismatched sizes hangs NCCL): shards are padded to the
    widest on the wire and trimmed with the host-side counts; collation goes to rank 0 only."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_uneven, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=100)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert ok
