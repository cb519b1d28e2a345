"""world_size-2 gloo test (CPU) of the N > 1 host logic: dictionary sharding by frames, global index bases, the
all-gather layout bench.py feeds to ss_topk_merge_dev ([rank][nq][k]), and the lexicographic (distance, index) merge
rule. Per-shard top-k come from the oracle here (no GPU on this box); the GPU run of the same path is
bench.py --gpus N."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def merge_lists(idx, dst, k):
    """numpy statement of ss_topk_merge_dev: idx/dst [nlists, nq, k] -> [nq, k], (distance, index) lexicographic."""
    nl, nq, _ = idx.shape
    oi = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
    od = np.full((nq, k), np.inf)
    for q in range(nq):
        cand = [(dst[l, q, s], int(idx[l, q, s])) for l in range(nl) for s in range(k) if idx[l, q, s] != 0xFFFFFFFF and dst[l, q, s] == dst[l, q, s]]
        cand.sort()
        for s, (d, i) in enumerate(cand[:k]):
            od[q, s], oi[q, s] = d, i
    return oi, od


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle import oracle as O
    from soundsym_b200 import synth
    d, doff = synth.segments(301, 13, seed=7)
    q, qoff = synth.segments(23, 13, seed=8)
    # duplicate a segment across the shard boundary so that a tie must resolve to the lower GLOBAL index
    k = 3
    cuts = bench.shard_bounds(doff, world)
    s0, s1 = cuts[rank], cuts[rank + 1]
    base = int(doff[s0])
    li, ld = O.dtw_topk(d[base:int(doff[s1])], doff[s0:s1 + 1] - doff[s0], q, qoff, 13, k)
    li = np.where(li == 0xFFFFFFFF, li, li + np.uint32(s0)).astype(np.uint32)  # index_base
    gi = [torch.empty((len(qoff) - 1, k), dtype=torch.int64) for _ in range(world)]
    gd = [torch.empty((len(qoff) - 1, k), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gi, torch.from_numpy(li.astype(np.int64)))
    dist.all_gather(gd, torch.from_numpy(ld))
    mi, md = merge_lists(np.stack([t.numpy() for t in gi]).astype(np.uint32), np.stack([t.numpy() for t in gd]), k)
    wi, wd = O.dtw_topk(d, doff, q, qoff, 13, k)
    ok = bool(np.array_equal(mi, wi) and np.array_equal(md, wd) and cuts[0] == 0 and cuts[-1] == 301 and 0 < cuts[1] < 301)
    frames = [int(doff[cuts[r + 1]] - doff[cuts[r]]) for r in range(world)]
    ok = ok and abs(frames[0] - frames[1]) <= 64  # balanced by frames, not by segment count
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_gather_merge():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret[0] is True and ret[1] is True


def test_merge_rule_ties_and_empty_slots():
    idx = np.array([[[5, 9, 0xFFFFFFFF]], [[2, 7, 8]]], dtype=np.uint32)
    dst = np.array([[[1.0, 2.0, np.inf]], [[1.0, 2.0, 2.5]]])
    oi, od = merge_lists(idx, dst, 3)
    assert list(oi[0]) == [2, 5, 7] and list(od[0]) == [1.0, 1.0, 2.0]


def test_shard_bounds_cover_and_balance():
    sys.path.insert(0, ROOT)
    import bench
    from soundsym_b200 import synth
    _, doff = synth.segments(1000, 13, seed=1)
    for n in (1, 2, 4, 8):
        cuts = bench.shard_bounds(doff, n)
        assert cuts[0] == 0 and cuts[-1] == 1000 and all(b > a for a, b in zip(cuts, cuts[1:]))
        fr = [int(doff[b] - doff[a]) for a, b in zip(cuts, cuts[1:])]
        assert max(fr) - min(fr) <= 64
