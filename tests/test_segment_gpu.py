"""GPU parity tests of the segmentation subsystem and of the three example flows."""
import numpy as np
import pytest

from oracle import oracle as O
from soundsym_b200 import api, synth
from soundsym_b200._lib import SS_COSINE_REF, SS_DTW, SoundsymError

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return api.Context(0)


def model_of(g):
    return (g["gmm_means"], g["gmm_covs"], g["gmm_weights"])


def test_symbols_section71_bit_exact(ctx, section71):
    sym, post = ctx.symbols(section71["mfcc"], model_of(section71), want_posteriors=True)
    assert np.array_equal(sym, section71["symbols"])  # boundary of parity: given the model, symbols are exact
    z, _, _ = O.standardize(section71["mfcc"])
    ref = O.gmm_posteriors(z, *model_of(section71))
    assert np.allclose(post, ref, rtol=1e-9, atol=1e-300)


def test_symbols_other_sound_uses_its_own_statistics(ctx, section71, sample_excerpt):
    """partition_other re-fits the Standardizer on the TARGET's MFCCs (src/lib.rs:57-58, reconstruction.rs:71-77);
    the excerpt has ~1 800 digitally silent frames."""
    sym = ctx.symbols(sample_excerpt["mfcc"], model_of(section71))
    assert np.array_equal(sym, sample_excerpt["symbols"])


def test_votes_and_splits_section71(ctx, section71):
    for depth, thr in ((3, 4), (4, 3), (5, 4)):
        votes, lens = ctx.vote_split(section71["symbols"], depth, thr)
        assert np.array_equal(votes, section71["votes_d%d" % depth])
        assert np.array_equal(lens, section71["splits_d%dt%d" % (depth, thr)])
        assert int(lens.sum()) == 1978 * 256


def test_vote_split_random_strings_and_edges(ctx):
    rng = np.random.default_rng(3)
    for n, alpha, depth, thr in ((5000, 26, 5, 4), (777, 4, 3, 2), (64, 2, 7, 1), (300, 26, 1, 1), (10, 3, 2, 0)):
        sym = (65 + rng.integers(0, alpha, n)).astype(np.uint8)
        votes, lens = ctx.vote_split(sym, depth, thr)
        ov = O.cast_votes(sym, depth)
        assert np.array_equal(votes, ov), (n, alpha, depth)
        assert np.array_equal(lens, O.split(ov, n, thr) * np.uint64(256))
    votes, lens = ctx.vote_split(np.array([65, 66], dtype=np.uint8), 3, 1)  # shorter than the window
    assert list(votes) == [0, 0, 0] and list(lens) == [512]
    votes, lens = ctx.vote_split(np.zeros(0, dtype=np.uint8), 3, 1)
    assert len(lens) == 0
    with pytest.raises(SoundsymError):
        ctx.vote_split(np.zeros(10, dtype=np.uint8), 8, 1)


def test_partition_errors_and_flow(ctx, section71):
    with pytest.raises(SoundsymError) as e:  # CosError("Must first train model"), src/lib.rs:140-142
        ctx.partition(section71["mfcc"], None, 3, 4)
    assert e.value.code == -4 and "Must first train model" in e.value.message
    with pytest.raises(SoundsymError) as e:
        ctx.partition(section71["mfcc"][:1], model_of(section71), 3, 4)
    assert e.value.code == -6
    lens = ctx.partition(section71["mfcc"], model_of(section71), 4, 3)  # examples/partition.rs defaults
    assert np.array_equal(lens, section71["splits_d4t3"])


def test_partition_long_synthetic(ctx, section71):
    """5 minutes of synthetic audio through MFCC -> symbols -> votes -> splits, GPU vs oracle end to end."""
    s = synth.audio(300.0, seed=42)
    m = ctx.mfcc(s)
    om = O.mfcc(s)
    assert np.all(np.abs(m - om) <= 1e-9 * np.maximum(np.abs(om), 1.0))
    z, _, _ = O.standardize(om)
    model = O.gmm_train(z, seed=1)
    lens = ctx.partition(m, model, 3, 4)
    olens, osym, _ = O.partition(om, model, 3, 4)
    sym = ctx.symbols(m, model)
    assert np.mean(sym != osym) < 1e-4  # MFCC inputs differ by ~1e-13: a symbol may flip only at a posterior tie
    if np.array_equal(sym, osym):
        assert np.array_equal(lens, olens)


def test_reconstruction_flow_like_the_example(ctx, section71, sample_excerpt):
    """examples/reconstruction.rs with tests/Section_7_1.wav as source and the sample.wav excerpt as target:
    Sound::from_samples -> Partitioner(threshold 4, depth 3).partition -> SoundDictionary::from_segments ->
    partition target with the source's model -> clone_from_dictionary -> to_sound."""
    src = api.Sound.from_samples(O.decode_pcm(section71["pcm"].astype(np.int32), 16), 44100.0, ctx=ctx)
    part = api.Partitioner(src, ctx).set_threshold(4).set_depth(3)
    part.train(model_of(section71))
    splits = part.partition()
    assert np.array_equal(splits, section71["splits_d3t4"])
    dictionary = api.SoundDictionary.from_segments(src, splits, ctx, SS_COSINE_REF)
    tgt = api.Sound.from_samples(O.decode_pcm(sample_excerpt["pcm"], 24), 44100.0, ctx=ctx)
    part.sound = tgt
    tsplits = part.partition()
    assert np.array_equal(tsplits, sample_excerpt["splits_d3t4"])
    segs, spos, fpos = [], 0, 0
    for sp in tsplits:  # reconstruction.rs:77-81
        sp = int(sp)
        segs.append(api.Sound(tgt.samples()[spos:spos + sp], 44100.0, tgt.mfcc_arrays()[fpos:fpos + sp // 256], None,
                              tgt.mfcc_arrays()[fpos:fpos + sp // 256].mean(axis=0) if sp // 256 else np.full(12, np.nan), None, ctx))
        spos += sp
        fpos += sp // 256
    seq = api.SoundSequence(segs, ctx)
    idx, dist = dictionary.match_indices(segs)
    assert np.array_equal(idx[:, 0], sample_excerpt["cos_idx"])  # MFCCs differ ~1e-13 from the oracle's: ties aside, same argmin
    out = seq.clone_from_dictionary(dictionary).to_sound()
    assert len(out.samples()) == int(sample_excerpt["resynth_len"])
    assert np.array_equal(out.samples()[:65536], sample_excerpt["resynth_head"])
    # matcher.rs flow: silent queries (max_power < 0.03) are gated out by the caller, the rest matched
    loud = [s for s, p in zip(segs, sample_excerpt["q_max_power"]) if p >= 0.03]
    assert len(loud) == int((sample_excerpt["q_max_power"] >= 0.03).sum())
    m = dictionary.match_sound(loud[0])
    assert m is dictionary.sounds[int(sample_excerpt["cos_idx"][list(sample_excerpt["q_max_power"] >= 0.03).index(True)])]
    # the DTW extension on the same data
    d2 = api.SoundDictionary.from_segments(src, splits, ctx, SS_DTW)
    di, dd = d2.match_indices(segs, k=4)
    assert np.array_equal(di, sample_excerpt["dtw_idx"])
    assert np.allclose(dd, sample_excerpt["dtw_dist"], rtol=1e-9)


def test_gmm_train_on_device_matches_seeded_oracle(ctx, section71):
    """train_model (src/lib.rs:44-54) on the GPU vs the oracle's seeded EM: same initial rows (same mt19937_64 draw), so
    the models agree to summation-order noise; the reference itself is randomly seeded, so this is the strongest parity
    that exists for training."""
    m = section71["mfcc"]
    means, covs, weights = ctx.gmm_train(m, 26, 5, 0.1, seed=0)
    assert np.allclose(means, section71["gmm_means"], rtol=1e-8, atol=1e-10)
    assert np.allclose(covs, section71["gmm_covs"], rtol=1e-8, atol=1e-10)
    assert np.allclose(weights, section71["gmm_weights"], rtol=1e-9) and abs(weights.sum() - 1.0) < 1e-12
    sym = ctx.symbols(m, (means, covs, weights))
    assert np.mean(sym != section71["symbols"]) < 2e-3  # a frame may flip only where two posteriors tie to ~1e-9
    z, _, _ = O.standardize(m)
    for seed, iters in ((3, 1), (7, 0)):
        gm = ctx.gmm_train(m, 26, iters, 0.1, seed=seed)
        om = O.gmm_train(z, 26, iters, 0.1, seed=seed)
        for a, b in zip(gm, om):
            assert np.allclose(a, b, rtol=1e-8, atol=1e-10)
    with pytest.raises(SoundsymError) as e:
        ctx.gmm_train(m[:10], 26)
    assert e.value.code == -6
    p = api.Partitioner(api.Sound(np.zeros(0), 44100.0, m, 0.0, m.mean(axis=0), None, ctx), ctx).set_depth(3).set_threshold(4)
    p.train(seed=0)  # Partitioner::train without a supplied model
    lens = p.partition()
    assert int(lens.sum()) == 1978 * 256 and len(lens) > 50


def test_config5_reconstruction_synthetic_end_to_end(ctx):
    """config 5 in miniature: synthetic 44.1 kHz audio, source = first 40 s, target = last 40 s; MFCC -> train on the GPU ->
    partition source -> dictionary -> partition target with the source's model -> clone_from_dictionary -> to_sound,
    every stage compared with the oracle run on the same model."""
    audio = synth.audio(80.0, seed=42)
    src = api.Sound.from_samples(audio[: 40 * 44100], 44100.0, ctx=ctx)
    tgt = api.Sound.from_samples(audio[40 * 44100:], 44100.0, ctx=ctx)
    osrc, otgt = O.mfcc(src.samples()), O.mfcc(tgt.samples())
    assert np.all(np.abs(src.mfcc_arrays() - osrc) <= 1e-9 * np.maximum(np.abs(osrc), 1.0))
    part = api.Partitioner(src, ctx).set_threshold(4).set_depth(3)  # examples/reconstruction.rs:43-45
    part.train(seed=3)
    model = part.model
    splits = part.partition()
    osplits, osym, _ = O.partition(src.mfcc_arrays(), model, 3, 4)  # oracle on the SAME rows and model
    assert np.array_equal(ctx.symbols(src.mfcc_arrays(), model), osym)
    assert np.array_equal(splits, osplits) and int(splits.sum()) == src.num_frames() * 256
    for mode in (SS_COSINE_REF, SS_DTW):
        dictionary = api.SoundDictionary.from_segments(src, splits, ctx, mode)
        part.sound = tgt
        tsplits = part.partition()
        otsplits, _, _ = O.partition(tgt.mfcc_arrays(), model, 3, 4)
        assert np.array_equal(tsplits, otsplits)
        segs, spos, fpos = [], 0, 0
        for sp in tsplits:
            sp = int(sp)
            m = tgt.mfcc_arrays()[fpos:fpos + sp // 256]
            segs.append(api.Sound(tgt.samples()[spos:spos + sp], 44100.0, m, None, m.mean(axis=0), None, ctx))
            spos += sp
            fpos += sp // 256
        idx, dist = dictionary.match_indices(segs)
        # the oracle's matcher on the same MFCC rows
        doff = np.zeros(len(splits) + 1, dtype=np.uint64)
        doff[1:] = np.cumsum(splits // np.uint64(256))
        qoff = np.zeros(len(tsplits) + 1, dtype=np.uint64)
        qoff[1:] = np.cumsum(tsplits // np.uint64(256))
        dm, qm = src.mfcc_arrays()[: int(doff[-1])], tgt.mfcc_arrays()[: int(qoff[-1])]
        if mode == SS_COSINE_REF:
            oi, od = O.cosine_match(dm, doff, qm, qoff, 12)
            assert np.array_equal(idx[:, 0], oi) and np.array_equal(dist[:, 0], od)
        else:
            oi, od = O.dtw_topk(dm, doff, qm, qoff, 12, 1)
            assert np.array_equal(idx, oi) and np.allclose(dist, od, rtol=1e-12, atol=0)
        out = api.SoundSequence(segs, ctx).clone_from_dictionary(dictionary).to_sound()
        soff = np.zeros(len(splits) + 1, dtype=np.uint64)
        soff[1:] = np.cumsum(splits)
        ref = O.resynth(src.samples(), soff, idx[:, 0], tsplits)
        assert np.array_equal(out.samples(), ref) and out.num_frames() == O.frame_count(len(ref))
