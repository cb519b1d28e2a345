"""GPU parity at the sizes the headline is measured on (BASELINE.json configs 3 and 4), and an on-hardware measurement of
the tensor-core scan's error against the slack its certification assumes.

  config 3 (10 000 x 1 000, C = 13), k in {1, 8}: EVERY index and distance vs the f64 oracle.
  config 4 (100 000 x 10 000): 512 random queries x the full dictionary, as one shard and as 8 shards
            (index_base + ss_topk_merge_dev), plus the whole batch through invariants (certified, work count,
            single-shard == 8-shard for all 10 000 queries).
  scan error: ss_dict_debug_tc_scan returns the raw scan distance of every pair; the host rebuilds the fp16 operands
            (same mean frame, same rounding) and compares with the f64 DTW of those ROUNDED frames. The certification
            (exact.cu scan_lower_bound, bound_mode 1) assumes  scan <= DTW(rounded) + E32,  E32 = 2e-5 (|a|^2 + |b|^2):
            the test asserts the measured excess is <= E32 / 4 and prints both sides.

Spec: oracle/ASSUMPTIONS.h A8 (DTW), tie rule /root/reference/src/sound.rs:361-366 (first minimum wins -> (distance, index)).
"""
import numpy as np
import pytest

from oracle import oracle as O
from soundsym_b200 import api, synth
from soundsym_b200._lib import SS_DTW

pytestmark = pytest.mark.gpu

C = 13


@pytest.fixture(scope="module")
def ctx():
    return api.Context(0)


@pytest.fixture(scope="module")
def config4():
    d, doff = synth.segments(100000, C, seed=1234)
    q, qoff = synth.segments(10000, C, seed=5678)
    return d, doff, q, qoff


def subset(q, qoff, ids):
    lens = (qoff[1:] - qoff[:-1]).astype(np.int64)[ids]
    off = np.zeros(len(ids) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    rows = np.concatenate([np.arange(int(qoff[i]), int(qoff[i + 1])) for i in ids]) if len(ids) else np.zeros(0, dtype=np.int64)
    return np.ascontiguousarray(q[rows]), off


def check(idx, dist, oidx, odist):
    assert np.array_equal(idx, oidx), "indices differ from the f64 oracle at %s" % (np.argwhere(idx != oidx)[:5].tolist(),)
    assert np.allclose(dist, odist, rtol=1e-12, atol=0)  # the refine kernel reproduces the oracle's arithmetic (bar: 1e-4)


def shard_bounds(doff, n):
    total = int(doff[-1])
    return [0] + [int(np.searchsorted(doff, total * r // n, side="left")) for r in range(1, n)] + [len(doff) - 1]


@pytest.mark.parametrize("first_stage", [0, 1, 2, 3])
def test_config3_every_query_vs_oracle(ctx, first_stage):
    """first_stage: 0 = packed-half tensor-core scan (the default), 1 = fp32-DP tensor-core scan, 2 = fp32 CUDA-core scan,
    3 = packed-half scan without its second chance - every filter is followed by the f64 refine + certification, so all of
    them must return the oracle's answer"""
    d, doff = synth.segments(10000, C, seed=1234)
    q, qoff = synth.segments(1000, C, seed=5678)
    O.set_threads(O.hardware_threads())
    dev = api.DeviceDictionary(ctx, d, doff)
    dev.set_scan(first_stage)
    for k in (1, 8):
        idx, dist = dev.match(q, qoff, SS_DTW, k)
        oidx, odist = O.dtw_topk(d, doff, q, qoff, C, k)
        check(idx, dist, oidx, odist)
        assert dev.last_uncertified == 0 and dev.last_exhaustive == 0
        assert dev.last_work == int(doff[-1]) * int(qoff[-1])
        if first_stage == 3:
            assert dev.last_tc_fallback <= 50  # the packed-half filter's merged list certifies all but a few per cent of the queries
        if first_stage == 0:
            assert dev.last_tc_fallback <= 2   # and the second chance (union of the per-slice lists) nearly all of the rest
        print("\nconfig 3, first stage %d, k=%d: %d queries left the first stage" % (first_stage, k, dev.last_tc_fallback))


def test_config4_sampled_queries_one_and_eight_shards_vs_oracle(ctx, config4):
    import torch
    d, doff, q, qoff = config4
    nq = len(qoff) - 1
    O.set_threads(O.hardware_threads())
    ids = np.sort(np.random.default_rng(99).choice(nq, size=512, replace=False))
    sq, sqoff = subset(q, qoff, ids)
    k = 4
    oidx, odist = O.dtw_topk(d, doff, sq, sqoff, C, k)
    # one shard, the whole batch of 10 000 (the bench's step), checked on the sample
    whole = api.DeviceDictionary(ctx, d, doff)
    widx, wdist = whole.match(q, qoff, SS_DTW, k)
    assert whole.last_uncertified == 0 and whole.last_exhaustive == 0
    assert whole.last_work == int(doff[-1]) * int(qoff[-1])
    check(widx[ids], wdist[ids], oidx, odist)
    del whole
    # eight shards with global index bases, merged on the device: identical for ALL queries, oracle-checked on the sample
    cuts = shard_bounds(doff, 8)
    li, ld = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        sh = api.DeviceDictionary(ctx, d, doff[a:b + 1], C, index_base=a)
        i, dd = sh.match(q, qoff, SS_DTW, k)
        assert sh.last_uncertified == 0
        li.append(i)
        ld.append(dd)
        del sh
    gi = torch.from_numpy(np.stack(li).astype(np.int64)).to(torch.int32).cuda()  # bit pattern of the u32 indices
    gd = torch.from_numpy(np.stack(ld)).cuda()
    oi = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    od = torch.empty((nq, k), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.check(ctx.lib.ss_topk_merge_dev(ctx.h, gi.data_ptr(), gd.data_ptr(), 8, nq, k, oi.data_ptr(), od.data_ptr()))
    ctx.sync()
    midx, mdist = oi.cpu().numpy().view(np.uint32), od.cpu().numpy()
    assert np.array_equal(midx, widx) and np.array_equal(mdist, wdist)
    check(midx[ids], mdist[ids], oidx, odist)


def rounded_operands(x, mu):
    """the scan's fp16 operands as f64: fp16(fp32(x - mu)), exactly k_tc_dict_tiles / k_tc_query_tiles' conversion"""
    return (x - mu).astype(np.float32).astype(np.float16).astype(np.float64)


@pytest.mark.parametrize("nd,nsample", [(10000, 256), (100000, 128)])
def test_tensor_core_scan_error_vs_certification_slack(ctx, config4, nd, nsample):
    d, doff, q, qoff = config4
    if nd != 100000:
        d, doff = synth.segments(nd, C, seed=1234)
    nq = len(qoff) - 1
    O.set_threads(O.hardware_threads())
    ids = np.sort(np.random.default_rng(7).choice(nq, size=nsample, replace=False))
    sq, sqoff = subset(q, qoff, ids)
    dev = api.DeviceDictionary(ctx, d, doff)
    scan, mu, scale = dev.debug_tc_scan(sq, sqoff)
    assert scan.shape == (nsample, nd) and np.all(np.isfinite(scan))
    dr, qr = rounded_operands(d, mu[:C]), rounded_operands(sq, mu[:C])
    ref = O.dtw_matrix(dr, doff, qr, sqoff, C)  # f64 DTW of the ROUNDED frames: what the scan approximates
    # per-pair slack of scan_lower_bound's E32 term, with the norms it uses: the query's max row norm, the dictionary's max
    na = np.array([np.max(np.sum(qr[int(sqoff[i]):int(sqoff[i + 1])] ** 2, axis=1)) for i in range(nsample)])
    nb = float(np.max(np.sum(dr ** 2, axis=1)))
    e32 = 2e-5 * (na + nb)
    err = scan.astype(np.float64) - ref
    over = np.max(err / e32[:, None])     # the direction the proof needs: scan must not EXCEED the rounded-frame DTW by more than E32
    under = np.max(-err / (na[:, None] * 2.0 ** -10 + e32[:, None]))  # informational: |a|^2 rides rounded DOWN (<= 2^-10 |a|^2 per cell)
    print("\nscan error over %d pairs (nd=%d): max (scan - DTW_rounded) = %.3e = %.4f E32; max (DTW_rounded - scan) = %.3e = %.4f of (2^-10 |a|^2 + E32); "
          "E32 in [%.3e, %.3e], DTW_rounded in [%.3e, %.3e]" % (err.size, nd, err.max(), over, (-err).max(), under, e32.min(), e32.max(), ref.min(), ref.max()))
    assert over <= 0.25, "tensor-core scan exceeds the rounded-frame DTW by %.3f of the E32 slack the certification assumes" % over
    assert under <= 1.0, "scan is lower than the rounded-frame DTW by more than the rd(|a|^2) budget"
    # and the end result on the same sample is the oracle's
    idx, dist = dev.match(sq, sqoff, SS_DTW, 1)
    oidx, odist = O.dtw_topk(d, doff, sq, sqoff, C, 1)
    check(idx, dist, oidx, odist)
    assert dev.last_uncertified == 0


@pytest.mark.parametrize("nd,nsample", [(10000, 256), (100000, 128)])
def test_packed_half_scan_is_a_lower_bound_of_the_rounded_frame_dtw(ctx, config4, nd, nsample):
    """The certification of the packed-half scan (exact.cu scan_lower_bound, bound_mode 2) rests on
        (scan - eta) (1 + 2^-11)^-(Lq + Ld + 2)  <=  DTW of the fp16-ROUNDED frames
    (every rounding of the F16 accumulator and of the half2 running sums is round-to-nearest, one per cost and one per cell on
    the path; a path sum beyond the fp16 range reads +inf). Measured here on every pair of a sample: the property must hold
    with room to spare, and the scan must stay tight enough to be a useful filter (printed)."""
    d, doff, q, qoff = config4
    if nd != 100000:
        d, doff = synth.segments(nd, C, seed=1234)
    nq = len(qoff) - 1
    O.set_threads(O.hardware_threads())
    ids = np.sort(np.random.default_rng(11).choice(nq, size=nsample, replace=False))
    sq, sqoff = subset(q, qoff, ids)
    dev = api.DeviceDictionary(ctx, d, doff)
    scan, mu, scale, S = dev.debug_h2_scan(sq, sqoff)
    assert scan.shape == (nsample, nd) and not np.any(np.isnan(scan))
    dr, qr = rounded_operands(d, mu[:C]), rounded_operands(sq, mu[:C])
    ref = O.dtw_matrix(dr, doff, qr, sqoff, C)
    lq = (sqoff[1:] - sqoff[:-1]).astype(np.float64)[:, None]
    ld = (doff[1:] - doff[:-1]).astype(np.float64)[None, :]
    u = 2.0 ** -11
    eta = (13.0 * float(np.max(np.abs(dr))) + 4.0) * 2.0 ** -24 / S
    fin = np.isfinite(scan)
    lower = (scan.astype(np.float64) - eta) * (1.0 + u) ** -(lq + ld + 2.0)
    ratio = np.max(lower[fin] / ref[fin])
    rel = scan.astype(np.float64)[fin] / ref[fin] - 1.0
    budget = ((1.0 + u) ** (lq + ld + 2.0) - 1.0 + 0 * ref)[fin]
    print("\npacked-half scan over %d pairs (nd=%d, S=%g): scan/DTW_rounded - 1 in [%.3e, %.3e] (mean %.2e); worst use of the rounding budget "
          "(1+u)^(Lq+Ld+2)-1: %.3f; max lower/ref = %.6f; %d pairs read +inf" % (scan.size, nd, S, rel.min(), rel.max(), rel.mean(),
                                                                                 float(np.max(rel / budget)), ratio, int((~fin).sum())))
    assert ratio <= 1.0, "the packed-half scan's certified lower bound exceeds the rounded-frame DTW"
    assert np.max(rel / budget) <= 0.5, "roundings use more than half of the worst-case budget: the bound has no margin"
    assert rel.min() >= -0.05  # the filter stays within 5 % below (|a|^2 rides rounded down; roundings cancel)
    # a pair may only read +inf if its path sum really is beyond the fp16 range
    if (~fin).any():
        assert np.all(ref[~fin] * (lq + ld + 0 * ref)[~fin] * S * (1.0 + u) ** (lq + ld + 2.0 + 0 * ref)[~fin] >= 65504.0 * 0.999)
    idx, dist = dev.match(sq, sqoff, SS_DTW, 1)
    oidx, odist = O.dtw_topk(d, doff, sq, sqoff, C, 1)
    check(idx, dist, oidx, odist)
    assert dev.last_uncertified == 0
