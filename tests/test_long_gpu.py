"""GPU parity of the DTW matcher on sequences of MORE than 32 frames (configs 1 and 5 of BASELINE.json: real segments reach
83 / 383 frames): the packed-half tensor-core scan cuts long dictionary segments into 32-column strips and streams long
query groups through a ring of A-block rows (dtw_h2.cu, k_dtw_scan_h2_long); its candidate keys are per-pair lower bounds
of the rounded-frame distance (exact.cu scan_lower_bound, bound_mode 3).

  * every index and distance vs the f64 oracle, for dictionaries / queries that mix short and long sequences, with the
    tensor-core path asserted to have been the one that ran (last_scan_kind == 1);
  * the property the certification rests on, measured on every pair:  key <= DTW(fp16-rounded frames);
  * config-1 fixtures (reference flow examples/matcher.rs:18-56 on the repo's two WAV files).

Spec: oracle/ASSUMPTIONS.h A8 (DTW), tie rule /root/reference/src/sound.rs:361-366.
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


def check(idx, dist, oidx, odist):
    assert np.array_equal(idx, oidx), "indices differ from the f64 oracle at %s" % (np.argwhere(idx != oidx)[:5].tolist(),)
    fin = np.isfinite(odist)
    assert np.array_equal(np.isfinite(dist), fin)
    assert np.allclose(dist[fin], odist[fin], rtol=1e-12, atol=0)


def mixed(nseg, seed, lens):
    """segments whose length ranges are drawn from `lens` = [(lmin, lmax, share), ..] and shuffled together"""
    parts, offs = [], [np.zeros(1, dtype=np.uint64)]
    rng = np.random.default_rng(seed)
    counts = [int(round(nseg * s)) for _, _, s in lens]
    counts[-1] = nseg - sum(counts[:-1])
    segs = []
    for (lo, hi, _), n, sd in zip(lens, counts, range(len(lens))):
        x, off = synth.segments(n, C, lmin=lo, lmax=hi, seed=seed * 31 + sd)
        segs += [x[int(off[i]):int(off[i + 1])] for i in range(n)]
    order = rng.permutation(len(segs))
    segs = [segs[i] for i in order]
    off = np.zeros(len(segs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in segs])
    return np.ascontiguousarray(np.concatenate(segs)), off


def rounded_operands(x, mu):
    return (x - mu).astype(np.float32).astype(np.float16).astype(np.float64)


@pytest.mark.parametrize("dl,ql", [
    ([(4, 32, 0.5), (33, 70, 0.3), (71, 200, 0.2)], [(4, 32, 0.6), (33, 120, 0.4)]),   # strips + streamed A blocks, all tile kinds
    ([(33, 64, 1.0)], [(4, 32, 1.0)]),                                                   # long segments, short queries
    ([(4, 32, 1.0)], [(33, 150, 1.0)]),                                                  # short segments, long queries (no strips)
    ([(120, 400, 1.0)], [(100, 400, 1.0)]),                                              # up to 13 strips, up to 50 A pieces
])
@pytest.mark.parametrize("k", [1, 4])
def test_long_sequences_every_query_vs_oracle_on_the_tensor_core_path(ctx, dl, ql, k):
    nd = 1500 if dl[-1][1] <= 200 else 300
    nq = 300 if ql[-1][1] <= 150 else 140
    d, doff = mixed(nd, 3, dl)
    q, qoff = mixed(nq, 5, ql)
    O.set_threads(O.hardware_threads())
    dev = api.DeviceDictionary(ctx, d, doff)
    idx, dist = dev.match(q, qoff, SS_DTW, k)
    assert dev.last_scan_kind == 1, "the packed-half tensor-core scan did not run (kind %d)" % dev.last_scan_kind
    oidx, odist = O.dtw_topk(d, doff, q, qoff, C, k)
    check(idx, dist, oidx, odist)
    assert dev.last_uncertified == 0
    assert dev.last_work == int(doff[-1]) * int(qoff[-1])
    print("\nlong sequences %s x %s, k=%d: %d of %d queries left the first pass, %d reached the exhaustive stage"
          % (dl, ql, k, dev.last_tc_fallback, nq, dev.last_exhaustive))


@pytest.mark.parametrize("dl,ql", [
    ([(4, 32, 0.4), (33, 100, 0.4), (101, 260, 0.2)], [(4, 32, 0.5), (33, 90, 0.5)]),
    ([(200, 383, 1.0)], [(200, 385, 1.0)]),
])
def test_strip_kernel_keys_are_lower_bounds_of_the_rounded_frame_dtw(ctx, dl, ql):
    """bound_mode 3: the key of a pair is (scan - eta)(1 + 2^-11)^-(Lq + Ld + 2), deflated by the scan itself, and must never
    exceed the f64 DTW of the fp16-rounded frames; a pair may only read +inf if its path sum leaves the fp16 range."""
    long_only = dl[0][0] >= 200
    d, doff = mixed(120 if long_only else 800, 7, dl)
    q, qoff = mixed(60 if long_only else 200, 9, ql)
    O.set_threads(O.hardware_threads())
    dev = api.DeviceDictionary(ctx, d, doff)
    key, mu, scale, S = dev.debug_h2_scan(q, qoff)
    nq, nd = len(qoff) - 1, len(doff) - 1
    assert key.shape == (nq, nd) and not np.any(np.isnan(key))
    dr, qr = rounded_operands(d, mu[:C]), rounded_operands(q, mu[:C])
    ref = O.dtw_matrix(dr, doff, qr, qoff, C)
    lq = (qoff[1:] - qoff[:-1]).astype(np.float64)[:, None]
    ld = (doff[1:] - doff[:-1]).astype(np.float64)[None, :]
    fin = np.isfinite(key)
    ratio = key.astype(np.float64)[fin] / ref[fin]
    u = 2.0 ** -11
    # what the raw scan was, relative to the rounded-frame DTW (undo the deflation; eta is negligible here)
    raw = ratio * (1.0 + u) ** ((lq + ld + 2.0 + 0 * ref)[fin])
    print("\nstrip kernel over %d pairs (S=%g): key / DTW_rounded in [%.4f, %.4f]; raw scan / DTW_rounded - 1 in [%.2e, %.2e]; %d pairs read +inf"
          % (key.size, S, ratio.min(), ratio.max(), raw.min() - 1.0, raw.max() - 1.0, int((~fin).sum())))
    assert ratio.max() <= 1.0, "a key exceeds the rounded-frame DTW: the certification of bound_mode 3 would not hold"
    assert ratio.min() >= 0.5  # still a filter
    if (~fin).any():
        assert np.all((ref * (lq + ld) * S * (1.0 + u) ** (lq + ld + 2.0))[~fin] >= 65504.0 * 0.999)
    idx, dist = dev.match(q, qoff, SS_DTW, 1)
    assert dev.last_scan_kind == 1
    oidx, odist = O.dtw_topk(d, doff, q, qoff, C, 1)
    check(idx, dist, oidx, odist)
    assert dev.last_uncertified == 0
