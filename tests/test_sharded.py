"""Multi-GPU matcher behind the C ABI: ss_shard_bounds, ss_comm_*, ss_dict_match_sharded,
ss_dict_create_sharded / ss_sharded_dict_match. The result must equal the single-GPU result bit for bit, whatever the
number of shards (the tie rule of /root/reference/src/sound.rs:361-366 carried across shards by (distance, index))."""
import numpy as np
import pytest

from soundsym_b200 import api, synth
from soundsym_b200._lib import SS_COSINE_REF, SS_DTW


def test_shard_bounds_host_arithmetic():
    """no GPU: contiguous ranges that cover every segment once and are balanced by FRAMES (work ~ Ld)"""
    import bench
    d, doff = synth.segments(5000, 13, seed=7)
    for n in (1, 2, 3, 4, 8):
        cuts = api.shard_bounds(doff, n)
        assert cuts == bench.shard_bounds(doff, n)  # bench.py's own statement of the rule
        assert cuts[0] == 0 and cuts[-1] == 5000 and all(a <= b for a, b in zip(cuts, cuts[1:]))
        frames = [int(doff[b] - doff[a]) for a, b in zip(cuts, cuts[1:])]
        assert sum(frames) == int(doff[-1]) and max(frames) - min(frames) <= 2 * 32
    # an offset table that does not start at 0 (a shard of a larger table), more shards than segments, empty segments
    sub = doff[100:104]
    assert api.shard_bounds(sub, 2) in ([0, 1, 3], [0, 2, 3])
    cuts = api.shard_bounds(np.array([0, 5, 9], dtype=np.uint64), 8)
    assert cuts[0] == 0 and cuts[-1] == 2 and all(a <= b for a, b in zip(cuts, cuts[1:]))
    cuts = api.shard_bounds(np.array([0, 0, 0, 7, 7, 12], dtype=np.uint64), 3)
    assert cuts[0] == 0 and cuts[-1] == 5 and all(a <= b for a, b in zip(cuts, cuts[1:]))


@pytest.fixture(scope="module")
def ctx():
    return api.Context(0)


@pytest.mark.gpu
def test_single_rank_communicator_equals_plain_match(ctx):
    """nranks = 1 drives the whole rank-local path (sharded query upload + all-gather, exchange block, merge) on one GPU"""
    d, doff = synth.segments(3000, 13, seed=51)
    q, qoff = synth.segments(200, 13, seed=52)
    dev = api.DeviceDictionary(ctx, d, doff)
    comm = api.Comm(ctx, 1, 0, api.Comm.unique_id())
    for mode, k, targets in ((SS_DTW, 1, None), (SS_DTW, 4, None), (SS_COSINE_REF, 1, None), (SS_COSINE_REF, 1, np.linspace(-1e-4, 1e-4, 200))):
        i0, d0 = dev.match(q, qoff, mode, k, targets)
        i1, d1 = comm.match(dev, q, qoff, mode, k, targets)
        assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    # empty query batch and a batch with empty queries
    i1, d1 = comm.match(dev, np.zeros((0, 13)), np.zeros(1, dtype=np.uint64), SS_DTW, 2)
    assert i1.shape == (0, 2)
    off = np.array([0, 5, 5, 12], dtype=np.uint64)
    i0, d0 = dev.match(q[:12], off, SS_DTW, 2)
    i1, d1 = comm.match(dev, q[:12], off, SS_DTW, 2)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1) and np.all(i1[1] == 0xFFFFFFFF)
    comm.close()


@pytest.mark.gpu
def test_sharded_dictionary_over_all_gpus_equals_single_gpu(ctx):
    """ss_dict_create_sharded over every GPU of the box (1 on the single-GPU test box, N under gpurun --gpus N): one worker
    thread per GPU, NCCL all-gather of the per-shard top-k, merge; equal to the single-GPU answer for DTW (k = 1, 4) and
    for the reference's cosine matcher with targets."""
    import torch
    n = min(torch.cuda.device_count(), 8)
    ctxs = [ctx] + [api.Context(i) for i in range(1, n)]
    d, doff = synth.segments(6000, 13, seed=61)
    q, qoff = synth.segments(500, 13, seed=62)
    # a duplicate of segment 10 at the very end: with more than one shard the tie must resolve to the lower GLOBAL index
    seg = d[int(doff[10]):int(doff[11])]
    d = np.concatenate([d, seg])
    doff = np.concatenate([doff, [doff[-1] + np.uint64(len(seg))]]).astype(np.uint64)
    q = np.concatenate([q, seg])
    qoff = np.concatenate([qoff, [qoff[-1] + np.uint64(len(seg))]]).astype(np.uint64)
    whole = api.DeviceDictionary(ctx, d, doff)
    sharded = api.ShardedDictionary(ctxs, d, doff)
    assert len(sharded) == 6001
    targets = np.linspace(-1e-4, 1e-4, len(qoff) - 1)
    for mode, k, t in ((SS_DTW, 1, None), (SS_DTW, 4, None), (SS_COSINE_REF, 1, None), (SS_COSINE_REF, 1, targets)):
        i0, d0 = whole.match(q, qoff, mode, k, t)
        for _ in range(2):  # the second call reuses every workspace
            i1, d1 = sharded.match(q, qoff, mode, k, t)
            assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    i1, d1 = sharded.match(q, qoff, SS_DTW, 2)
    assert list(i1[-1]) == [10, 6000] and d1[-1, 0] == 0.0 and d1[-1, 1] == 0.0
    # fewer segments than GPUs: empty shards contribute nothing
    tiny = api.ShardedDictionary(ctxs, d[: int(doff[1])], doff[:2])
    i1, d1 = tiny.match(q, qoff, SS_DTW, 2)
    assert np.all(i1[:, 0] == 0) and np.all(i1[:, 1] == 0xFFFFFFFF)
    # errors surface through ctxs[0]
    with pytest.raises(api.SoundsymError):
        sharded.match(q, qoff, SS_DTW, 9)
    sharded.close(), tiny.close()
