"""CPU tests of the oracle itself: reference known answers, golden vectors, and C++-oracle vs numpy-twin agreement.
(-m "not gpu"; nothing here touches the product library.)"""
import json
import os

import numpy as np
import pytest

from oracle import numpy_twin as T
from oracle import oracle as O
from soundsym_b200 import synth


def test_reference_known_answers(golden_dir):
    kats = json.load(open(os.path.join(golden_dir, "kats.json")))
    for name, k in kats.items():
        v, e = np.asarray(k["value"], dtype=np.float64), np.asarray(k["expected"], dtype=np.float64)
        assert np.all(np.abs(v - e) <= k.get("tol", 0.0)), name


def test_decode_scale_matches_reference_rule():
    # src/sound.rs:118-120: s / (i32::MAX >> (32 - bits))
    pcm = np.array([0, 1, -1, 32767, -32768], dtype=np.int32)
    assert np.array_equal(O.decode_pcm(pcm, 16), pcm / 32767.0)
    assert O.decode_pcm(np.array([8388607], dtype=np.int32), 24)[0] == 1.0


def test_angular_distance_kat():
    # src/sound.rs:612-615
    v = [0.1, 0.4, 0.2, 0.8] + [0.0] * 8
    assert O.cosine_sim_angular(v, v) == 0.0
    # the x < -1 -> 1 quirk (src/sound.rs:65): opposite vectors with sim < -1 also give 0
    w = [-x for x in v]
    assert O.cosine_sim_angular(v, w) == 0.0


def test_frame_count_rule():
    # A1; the reference's own asserts (src/sound.rs:618-631, expects 5) are stale vs BIN=1024/HOP=256 -> 17
    assert O.frame_count(5120) == 17
    assert O.frame_count(1023) == 0 and O.frame_count(1024) == 1 and O.frame_count(1279) == 1 and O.frame_count(1280) == 2
    assert O.frame_count(0) == 0


def test_mfcc_golden_and_twin(section71):
    s = O.decode_pcm(section71["pcm"].astype(np.int32), int(section71["bits"]))
    m = O.mfcc(s, float(section71["sample_rate"]))
    assert m.shape == (1978, 12)
    assert np.array_equal(m, section71["mfcc"])
    m2 = T.mfcc(s, float(section71["sample_rate"]))
    assert np.max(np.abs(m - m2)) < 1e-10
    assert O.max_power(s) == float(section71["max_power"])
    assert abs(O.max_power(s) - T.max_power(s)) < 1e-14
    assert np.array_equal(O.mean_mfccs(m), section71["mean_mfccs"])


def test_mfcc_silence_is_finite():
    m = O.mfcc(np.zeros(5120))
    assert m.shape == (17, 12) and np.all(np.isfinite(m))
    assert np.allclose(m[:, 1:], 0.0, atol=1e-9) and np.allclose(m[:, 0], 2 * 12 * -10.0)
    assert np.all(np.isnan(O.mean_mfccs(np.zeros((0, 12)))))


def test_c13_variant_matches_twin():
    s = synth.audio(0.5, seed=7)
    assert np.max(np.abs(O.mfcc(s, ncoeffs=13) - T.mfcc(s, ncoeffs=13))) < 1e-10


def test_symbols_votes_splits_golden(section71):
    model = (section71["gmm_means"], section71["gmm_covs"], section71["gmm_weights"])
    sym = O.symbols(section71["mfcc"], *model)
    assert np.array_equal(sym, section71["symbols"])
    post = T.gmm_posteriors(T.standardize(section71["mfcc"]), *model)
    assert np.array_equal(sym, np.array([65 + T.max_index(r) for r in post], dtype=np.uint8))
    for depth, thr in ((3, 4), (4, 3), (5, 4)):
        v = O.cast_votes(sym, depth)
        assert np.array_equal(v, section71["votes_d%d" % depth])
        assert np.array_equal(v, T.cast_votes(sym, depth))
        sp = O.split(v, len(sym), thr)
        assert np.array_equal(sp * np.uint64(256), section71["splits_d%dt%d" % (depth, thr)])
        assert np.array_equal(sp, T.split(v, len(sym), thr))
        assert int(sp.sum()) == len(sym)  # chunks cover the whole string (src/lib.rs:137)


def test_votes_edge_cases():
    assert np.array_equal(O.cast_votes(np.zeros(0, dtype=np.uint8), 3), [0])
    assert np.array_equal(O.cast_votes(np.array([65, 66], dtype=np.uint8), 3), [0, 0, 0])  # shorter than the window
    assert len(O.split(np.zeros(1, dtype=np.uint32), 0, 3)) == 0
    sym = np.frombuffer(b"ABABABABCABABABABC" * 4, dtype=np.uint8)
    for depth in (1, 2, 3, 5):
        assert np.array_equal(O.cast_votes(sym, depth), T.cast_votes(sym, depth))


def test_max_index_rule():
    # src/sound.rs:486-495: starts at (0, 0.0), strict '>'
    assert O.lib().orc_max_index(np.array([0.0, 0.0]), 2) == 0
    assert O.lib().orc_max_index(np.array([np.nan, np.nan]), 2) == 0
    assert O.lib().orc_max_index(np.array([0.2, 0.5, 0.5]), 3) == 1
    assert O.lib().orc_max_index(np.array([-1.0, -0.5]), 2) == 0


def test_cosine_matcher_golden(section71, sample_excerpt):
    from tests.golden.make_golden import cut
    d, doff = cut(section71["mfcc"], section71["splits_d3t4"])
    q, qoff = cut(sample_excerpt["mfcc"], sample_excerpt["splits_d3t4"])
    idx, dist = O.cosine_match(d, doff, q, qoff, 12)
    assert np.array_equal(idx, sample_excerpt["cos_idx"]) and np.array_equal(dist, sample_excerpt["cos_dist"])
    # twin agreement on the similarity itself
    for qi in (0, 7, 50):
        a = q[int(qoff[qi]):int(qoff[qi + 1])]
        b = d[int(doff[3]):int(doff[4])]
        assert abs(O.cosine_sim(b, a) - T.cosine_sim(b, a)) <= 1e-12 * abs(T.cosine_sim(b, a)) + 1e-300


def test_cosine_first_minimum_wins_and_nan_never_wins():
    d = np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 0.0]])
    doff = np.array([0, 1, 2, 3], dtype=np.uint64)
    q = np.array([[1.0, 0.0]])
    idx, dist = O.cosine_match(d, doff, q, np.array([0, 1], dtype=np.uint64), 2)
    assert idx[0] == 0 and dist[0] == 0.0  # entries 0 and 1 tie; entry 2 gives NaN
    idx, dist = O.cosine_match(d[2:], np.array([0, 1], dtype=np.uint64), q, np.array([0, 1], dtype=np.uint64), 2)
    assert idx[0] == 0 and dist[0] == 2.0  # nothing < 2.0 -> index 0 (src/sound.rs:361-369)


def test_dtw_golden_and_twin(synthetic_small, section71, sample_excerpt):
    d, doff = synth.segments(600, 13, seed=1234)
    q, qoff = synth.segments(48, 13, seed=5678)
    assert float(d.sum()) == float(synthetic_small["dict_checksum"])
    idx, dist = O.dtw_topk(d, doff, q, qoff, 13, k=4)
    assert np.array_equal(idx, synthetic_small["dtw_idx"]) and np.array_equal(dist, synthetic_small["dtw_dist"])
    for qi, di in ((0, 0), (5, 77), (47, 599)):
        a = q[int(qoff[qi]):int(qoff[qi + 1])]
        b = d[int(doff[di]):int(doff[di + 1])]
        assert abs(O.dtw(a, b) - T.dtw(a, b)) <= 1e-12 * T.dtw(a, b)
        assert O.dtw(a, b) == O.dtw(b, a)  # symmetric step pattern
    assert O.dtw(d[:5], d[:5]) == 0.0
    assert O.dtw(d[:0], d[:5]) == np.inf


def test_dtw_topk_ties_resolve_to_lowest_index():
    d = np.tile(np.arange(6, dtype=np.float64).reshape(2, 3), (3, 1))  # three identical 2-frame segments
    doff = np.array([0, 2, 4, 6], dtype=np.uint64)
    idx, dist = O.dtw_topk(d, doff, d[:2], np.array([0, 2], dtype=np.uint64), 3, k=4)
    assert list(idx[0]) == [0, 1, 2, 0xFFFFFFFF] and list(dist[0][:3]) == [0.0, 0.0, 0.0] and dist[0][3] == np.inf


def test_resynth_pad_and_truncate(section71, sample_excerpt):
    ds = np.arange(10, dtype=np.float64)
    out = O.resynth(ds, np.array([0, 4, 10], dtype=np.uint64), np.array([1, 0, 0]), np.array([3, 6, 4], dtype=np.uint64))
    assert list(out) == [4, 5, 6, 0, 1, 2, 3, 0, 0, 0, 1, 2, 3]
    assert int(sample_excerpt["resynth_len"]) == int(sample_excerpt["splits_d3t4"].sum())


def test_gmm_train_is_seed_deterministic(section71):
    z, _, _ = O.standardize(section71["mfcc"])
    m1 = O.gmm_train(z, seed=0)
    assert np.array_equal(m1[0], section71["gmm_means"]) and np.array_equal(m1[1], section71["gmm_covs"])
    assert abs(m1[2].sum() - 1.0) < 1e-12
    with pytest.raises(ValueError):
        O.standardize(np.zeros((1, 12)))
