"""GPU parity of the at_distance(target != 1) consumers through BOTH host mirrors:
SoundSequence::morph_to (/root/reference/src/sound.rs:440-449), SoundSequence::from_distances (:405-417, a chain of
nq = 1 matches), the examples/matcher.rs:18-56 loop (one match_sound per file), and the max power that
SoundDictionary::add_segments' Sound::from_samples computes for every cut (:95, :330-343)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from soundsym_b200 import api
from soundsym_b200._lib import SS_COSINE_REF

from test_cpp_host_gpu import run, write_model, write_wav

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return api.Context(0)


def cut_offsets(splits, hop=256):
    frames = (np.asarray(splits, dtype=np.uint64) // np.uint64(hop)).astype(np.uint64)
    off = np.zeros(len(frames) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(frames)
    return off


def segments_of(ctx, samples, mfcc, splits):
    snd = api.Sound(samples, 44100.0, mfcc, None, mfcc.mean(axis=0), None, ctx)
    d = api.SoundDictionary.from_segments(snd, splits, ctx, SS_COSINE_REF)
    return d


def test_add_segments_cuts_carry_their_max_power_and_accept_push_samples(ctx, section71):
    s16 = O.decode_pcm(section71["pcm"].astype(np.int32), 16)
    splits = section71["splits_d3t4"]
    d = segments_of(ctx, s16, section71["mfcc"], splits)
    assert len(d.sounds) == len(splits)
    pos = 0
    for snd, sp in zip(d.sounds, splits):
        assert snd.max_power() == O.max_power(s16[pos:pos + int(sp)])  # bit-exact, as analyze_max_power on the cut
        pos += int(sp)
    # push_samples on a cut used to raise (max(None, x)); it must extend samples, MFCC rows and keep the louder power
    cutsnd = d.sounds[5]
    before, frames0 = cutsnd.max_power(), cutsnd.num_frames()
    extra = s16[100000:100000 + 4096]
    cutsnd.push_samples(extra)
    assert cutsnd.max_power() == max(before, O.max_power(cutsnd.samples()[frames0 * 256:]))
    assert cutsnd.num_frames() == frames0 + O.frame_count(len(cutsnd.samples()) - frames0 * 256)


def test_morph_to_and_from_distances_python_mirror_vs_oracle(ctx, section71, sample_excerpt):
    s16 = O.decode_pcm(section71["pcm"].astype(np.int32), 16)
    d = segments_of(ctx, s16, section71["mfcc"], section71["splits_d3t4"])
    tgt = O.decode_pcm(sample_excerpt["pcm"], 24)
    q = segments_of(ctx, tgt, sample_excerpt["mfcc"], sample_excerpt["splits_d3t4"])
    doff, qoff = cut_offsets(section71["splits_d3t4"]), cut_offsets(sample_excerpt["splits_d3t4"])
    dm, qm = section71["mfcc"][: int(doff[-1])], sample_excerpt["mfcc"][: int(qoff[-1])]
    seq = api.SoundSequence(q.sounds, ctx)
    # morph_to: sound i is replaced by the dictionary sound whose similarity to it is closest to distances[i]
    distances = np.linspace(-1e-4, 1e-4, len(q.sounds))
    morphed = seq.morph_to(distances, d)
    oidx, _ = O.cosine_match(dm, doff, qm, qoff, 12, distances)
    assert [d.sounds.index(s) for s in morphed.sounds()] == [int(i) for i in oidx]
    # zip semantics: the shorter of (sounds, distances) bounds the result
    assert len(seq.morph_to(distances[:7], d).sounds()) == 7
    # from_distances: a sequential chain, every step an nq = 1 at_distance from the previous RESULT
    steps = [5e-5, -2e-5, 1e-4, 0.0, -1e-4, 3e-5, 1.0, 2e-5]
    start = q.sounds[3]
    chain = api.SoundSequence.from_distances(steps, start, d)
    assert chain.sounds()[0] is start and len(chain.sounds()) == len(steps) + 1
    cur = (qm[int(qoff[3]):int(qoff[4])], np.array([0, qoff[4] - qoff[3]], dtype=np.uint64))
    for step, got in zip(steps, chain.sounds()[1:]):
        oi, _ = O.cosine_match(dm, doff, cur[0], cur[1], 12, np.array([step]))
        assert d.sounds.index(got) == int(oi[0])
        a, b = int(doff[oi[0]]), int(doff[oi[0] + 1])
        cur = (dm[a:b], np.array([0, b - a], dtype=np.uint64))
    # the sequence's consecutive distances are cosine_sim_angular of the mean MFCCs (src/sound.rs:392-396)
    means = np.stack([s.mean_mfccs() for s in chain.sounds()])
    want = np.array([O.cosine_sim_angular(means[i], means[i + 1]) for i in range(len(means) - 1)])
    assert np.allclose(chain.distances(), want, rtol=0, atol=1e-15, equal_nan=True)


def test_morph_to_and_from_distances_cpp_mirror_vs_oracle(tmp_path, section71, sample_excerpt):
    src, tgt, mdl = (str(tmp_path / n) for n in ("source.wav", "target.wav", "model.bin"))
    write_wav(src, section71["pcm"], 16)
    write_wav(tgt, sample_excerpt["pcm"], 24)
    write_model(mdl, section71)
    chain = 6
    rc, r = run(["-s", src, "-t", tgt, "-m", mdl, "--morph", "--chain", str(chain)])
    assert rc == 0
    doff, qoff = cut_offsets(section71["splits_d3t4"]), cut_offsets(sample_excerpt["splits_d3t4"])
    dm, qm = section71["mfcc"][: int(doff[-1])], sample_excerpt["mfcc"][: int(qoff[-1])]
    n = len(qoff) - 1
    distances = np.array([-1e-4 + 2e-4 * i / max(n - 1, 1) for i in range(n)])
    oidx, _ = O.cosine_match(dm, doff, qm, qoff, 12, distances)
    assert r["morph"] == [int(i) for i in oidx]
    steps = [(-1.0 if i % 2 else 1.0) * 5e-5 * (i + 1) / chain for i in range(chain)]
    cur = (qm[int(qoff[0]):int(qoff[1])], np.array([0, qoff[1] - qoff[0]], dtype=np.uint64))
    want = []
    for step in steps:
        oi, _ = O.cosine_match(dm, doff, cur[0], cur[1], 12, np.array([step]))
        want.append(int(oi[0]))
        a, b = int(doff[oi[0]]), int(doff[oi[0] + 1])
        cur = (dm[a:b], np.array([0, b - a], dtype=np.uint64))
    assert r["chain"] == want and len(r["chain_distances"]) == chain


def test_matcher_rs_loop_one_match_sound_per_file(tmp_path, section71):
    """examples/matcher.rs:18-56: per query file Sound::from_path, the 0.03 max-power gate, ONE match_sound (nq = 1), the
    matched samples padded / truncated to the query's length and written as (s * i16::MAX * 4^max_power) as i16.
    C++ mirror (the binary's --matcher-loop) vs the Python mirror vs the oracle."""
    ddir, qdir = tmp_path / "dict", tmp_path / "queries"
    ddir.mkdir(), qdir.mkdir()
    pcm = section71["pcm"]
    rng = np.random.default_rng(17)
    for i in range(24):  # dictionary: whole files of 0.1 .. 0.5 s
        a = int(rng.integers(0, len(pcm) - 30000))
        write_wav(str(ddir / ("d%02d.wav" % i)), pcm[a:a + int(rng.integers(4410, 22050))], 16)
    for i in range(9):
        a = int(rng.integers(0, len(pcm) - 30000))
        write_wav(str(qdir / ("p%d.wav" % i)), pcm[a:a + int(rng.integers(3000, 20000))], 16)
    write_wav(str(qdir / "hush.wav"), (rng.normal(size=6000) * 30).astype(np.int16), 16)  # max power ~1e-3: silent branch
    rc, r = run(["--matcher-loop", str(ddir), str(qdir)])
    assert rc == 0
    ctx = api.Context(0)
    d = api.SoundDictionary.from_path(str(ddir), ctx)
    names = sorted(os.listdir(qdir))
    matches, concat = [], []
    doff = np.zeros(len(d.sounds) + 1, dtype=np.uint64)
    doff[1:] = np.cumsum([s.num_frames() for s in d.sounds])
    dm = np.concatenate([s.mfcc_arrays() for s in d.sounds])
    for nme in names:
        ph = api.Sound.from_path(str(qdir / nme), ctx)
        if ph.max_power() < 0.03:
            concat.append(np.zeros(len(ph.samples()), dtype=np.int16))
            continue
        snd = d.match_sound(ph)  # nq = 1
        oi, _ = O.cosine_match(dm, doff, ph.mfcc_arrays(), np.array([0, ph.num_frames()], dtype=np.uint64), 12)
        assert d.sounds.index(snd) == int(oi[0])
        matches.append([ph.name, snd.name])
        s = np.zeros(len(ph.samples()))
        m = min(len(s), len(snd.samples()))
        s[:m] = snd.samples()[:m]
        concat.append(np.clip(np.trunc(s * 32767.0 * 4.0 ** ph.max_power()), -32768, 32767).astype(np.int16))
    concat = np.concatenate(concat)
    assert any(n == "hush.wav" for n in names) and len(matches) < len(names)
    assert r["matches"] == matches
    assert r["concat_len"] == len(concat) and r["concat_sum"] == int(concat.astype(np.int64).sum())
