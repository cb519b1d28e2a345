"""GPU parity tests of the MFCC subsystem against the oracle and the golden fixtures."""
import numpy as np
import pytest

from oracle import oracle as O
from soundsym_b200 import api, synth
from soundsym_b200._lib import SoundsymError

pytestmark = pytest.mark.gpu

# north-star bar: MFCC coefficients within 1e-4 relative. The kernel computes in f64, so the test holds it to
# |delta| <= 1e-9 * max(|ref|, 1) — five orders tighter; the slack covers a different FFT factorisation only.
MFCC_RTOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    return api.Context(0)


def assert_mfcc_close(got, ref):
    assert got.shape == ref.shape
    assert np.all(np.abs(got - ref) <= MFCC_RTOL * np.maximum(np.abs(ref), 1.0)), float(np.max(np.abs(got - ref)))


def test_decode_pcm_bit_exact(ctx, section71, sample_excerpt):
    s16 = ctx.decode_pcm(section71["pcm"].astype(np.int32), 16)
    assert np.array_equal(s16, O.decode_pcm(section71["pcm"].astype(np.int32), 16))
    s24 = ctx.decode_pcm(sample_excerpt["pcm"], 24)
    assert np.array_equal(s24, O.decode_pcm(sample_excerpt["pcm"], 24))
    assert ctx.decode_pcm(np.array([8388607], dtype=np.int32), 24)[0] == 1.0


def test_analyze_section71_vs_golden(ctx, section71):
    s = O.decode_pcm(section71["pcm"].astype(np.int32), 16)
    mfcc, mp, mean = ctx.analyze(s, float(section71["sample_rate"]), 12)
    assert mfcc.shape == (1978, 12)
    assert_mfcc_close(mfcc, section71["mfcc"])
    assert mp == float(section71["max_power"])  # bit-exact: same fold, separate multiply / add
    assert np.allclose(mean, section71["mean_mfccs"], rtol=1e-12, atol=1e-12)


def test_analyze_sample_excerpt_with_digital_silence(ctx, sample_excerpt):
    s = O.decode_pcm(sample_excerpt["pcm"], 24)
    mfcc, mp, _ = ctx.analyze(s, 44100.0, 12)
    assert np.all(np.isfinite(mfcc))  # silent frames hit the energy floor (A4), no -inf
    assert_mfcc_close(mfcc, sample_excerpt["mfcc"])
    assert mp == float(sample_excerpt["max_power"])


def test_c13_and_synthetic_audio(ctx):
    s = synth.audio(3.0, seed=5)
    for c in (12, 13):
        assert_mfcc_close(ctx.mfcc(s, 44100.0, c), O.mfcc(s, 44100.0, c))
    assert ctx.max_power(s) == O.max_power(s)


def test_other_sample_rates_change_the_mel_bank_layout(ctx):
    """The kernel reads its mel weights / DCT matrix / twiddles from tables laid out per lane: other sample rates give other
    band edges (wider half bands at 22.05 kHz, bins up to 342 instead of 229), other table shapes."""
    s = synth.audio(1.5, seed=9)
    for sr in (22050.0, 32000.0, 48000.0, 96000.0):
        for c in (12, 13):
            assert_mfcc_close(ctx.mfcc(s, sr, c), O.mfcc(s, sr, c))


def test_frame_edge_cases(ctx):
    assert ctx.mfcc(np.zeros(0)).shape == (0, 12)
    assert ctx.mfcc(np.zeros(1023)).shape == (0, 12)
    m = ctx.mfcc(np.zeros(5120))
    assert m.shape == (17, 12)  # the reference's stale asserts expect 5 (src/sound.rs:618-631); the code gives 17
    assert_mfcc_close(m, O.mfcc(np.zeros(5120)))  # silence: energy floor (A4) -> c0 = -240, others ~0
    rng = np.random.default_rng(0)
    for n in (1024, 1279, 1280, 4097):
        s = rng.uniform(-1, 1, n)
        assert_mfcc_close(ctx.mfcc(s), O.mfcc(s))
        assert ctx.max_power(s) == O.max_power(s)
    assert ctx.max_power(np.zeros(100)) == 0.0 and ctx.max_power(np.zeros(0)) == 0.0
    _, _, mean = ctx.analyze(np.zeros(10))
    assert np.all(np.isnan(mean))  # 0/0, as analyze_mean_mfccs on an empty sound
    with pytest.raises(SoundsymError):
        ctx.mfcc(np.zeros(4096), 8000.0)  # band edges beyond Nyquist bin range


def test_push_samples_equals_batch(ctx):
    """Sound::push_samples (src/sound.rs:145-164) re-analyses from initial_frames*HOP: same frames as one batch pass."""
    s = synth.audio(1.0, seed=9)
    whole = api.Sound.from_samples(s, 44100.0, ctx=ctx)
    inc = api.Sound.from_samples(s[:10000], 44100.0, ctx=ctx)
    inc.push_samples(s[10000:30000])
    inc.push_samples(s[30000:])
    assert inc.num_frames() == whole.num_frames()
    assert np.array_equal(inc.mfcc_arrays(), whole.mfcc_arrays())
    assert inc.max_power() == whole.max_power()
    z = api.Sound.from_samples(np.zeros(0), 44100.0, ctx=ctx)
    z.push_samples(np.zeros(4096 + 1024))
    assert z.num_frames() == 17 and len(z.samples()) == 5120


def test_mfcc_dev_large_matches_oracle_sample(ctx):
    """device-resident entry on a longer signal; spot-check frames against the oracle."""
    import torch
    s = synth.audio(30.0, seed=11)
    d = torch.from_numpy(s).cuda()
    frames = O.frame_count(len(s))
    out = torch.empty((frames, 12), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.check(ctx.lib.ss_mfcc_dev(ctx.h, d.data_ptr(), len(s), 44100.0, 12, out.data_ptr()))
    ctx.sync()
    got = out.cpu().numpy()
    for f0 in (0, 1234, frames - 40):
        ref = O.mfcc(s[f0 * 256:f0 * 256 + 1024 + 39 * 256])
        assert_mfcc_close(got[f0:f0 + 40], ref)


def test_resynth_and_sequence_distances(ctx, section71, sample_excerpt):
    ds = np.arange(10, dtype=np.float64)
    out = ctx.resynth(ds, np.array([0, 4, 10], dtype=np.uint64), np.array([1, 0, 0]), np.array([3, 6, 4], dtype=np.uint64))
    assert list(out) == [4, 5, 6, 0, 1, 2, 3, 0, 0, 0, 1, 2, 3]
    s71 = O.decode_pcm(section71["pcm"].astype(np.int32), 16)
    soff = np.zeros(len(section71["splits_d3t4"]) + 1, dtype=np.uint64)
    soff[1:] = np.cumsum(section71["splits_d3t4"])
    out = ctx.resynth(s71, soff, sample_excerpt["cos_idx"], sample_excerpt["splits_d3t4"])
    ref = O.resynth(s71, soff, sample_excerpt["cos_idx"], sample_excerpt["splits_d3t4"])
    assert np.array_equal(out, ref)
    assert np.array_equal(out[:65536], sample_excerpt["resynth_head"]) and len(out) == int(sample_excerpt["resynth_len"])
    rows = section71["mfcc"][:50]
    got = ctx.sequence_distances(rows)
    ref = np.array([O.cosine_sim_angular(rows[i], rows[i + 1]) for i in range(49)])
    assert np.allclose(got, ref, rtol=0, atol=1e-15)
    v = np.array([[0.1, 0.4, 0.2, 0.8] + [0.0] * 8] * 2)
    assert ctx.sequence_distances(v)[0] == 0.0  # the reference's KAT, src/sound.rs:612-615


def test_pcm_ingest_equals_decode_then_analyze(ctx, section71, sample_excerpt, tmp_path):
    """ss_sound_analyze_pcm: int16 / int24 PCM crosses PCIe, conversion on the device (src/sound.rs:118-120)."""
    s16 = O.decode_pcm(section71["pcm"].astype(np.int32), 16)
    samples, mfcc, mp, mean = ctx.analyze_pcm(section71["pcm"], 16)
    assert np.array_equal(samples, s16)
    m2, mp2, mean2 = ctx.analyze(s16)
    assert np.array_equal(mfcc, m2) and mp == mp2 and np.array_equal(mean, mean2)
    samples, mfcc, mp, _ = ctx.analyze_pcm(sample_excerpt["pcm"], 24)
    assert np.array_equal(samples, O.decode_pcm(sample_excerpt["pcm"], 24)) and mp == float(sample_excerpt["max_power"])
    assert_mfcc_close(mfcc, sample_excerpt["mfcc"])
    # Sound::from_path through a real WAV file
    import struct
    body = section71["pcm"].astype("<i2").tobytes()
    p = tmp_path / "s.wav"
    p.write_bytes(b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 44100, 88200, 2, 16) + b"data"
                  + struct.pack("<I", len(body)) + body)
    snd = api.Sound.from_path(str(p), ctx)
    assert snd.name == "s" and snd.num_frames() == 1978 and snd.max_power() == float(section71["max_power"])
    assert np.array_equal(snd.samples(), s16)
    with pytest.raises(SoundsymError):
        ctx.analyze_pcm(section71["pcm"], 8)


def test_chunked_ingest_equals_one_shot(ctx):
    """Inputs longer than one ingest chunk (4 Mi samples) go through the double-buffered path: the H2D copy of chunk k+1
    overlaps the conversion + MFCC of chunk k. Frames are independent, so the result must be bit-identical to analysing
    the same samples in one piece - checked on windows that straddle every chunk boundary and on both ends."""
    rng = np.random.default_rng(11)
    n = 9_500_000 + 777  # three chunks, ragged tail
    pcm = (rng.standard_normal(n) * 3000).astype(np.int16)
    samples, mfcc, mp, mean = ctx.analyze_pcm(pcm, 16)
    ref = pcm.astype(np.float64) / 32767.0
    assert np.array_equal(samples, ref)
    frames = (n - 1024) // 256 + 1
    assert mfcc.shape == (frames, 12)
    chunk_frames = (4 << 20) // 256
    for f0 in (0, chunk_frames - 40, 2 * chunk_frames - 40, frames - 80):
        sub = ref[f0 * 256: (f0 + 79) * 256 + 1024]  # 80 frames, far below the chunk size: one-shot path
        m_sub, _, _ = ctx.analyze(sub)
        assert np.array_equal(mfcc[f0:f0 + 80], m_sub), f0
    # f64 entry point takes the same path
    m64, mp64, mean64 = ctx.analyze(ref)
    assert np.array_equal(m64, mfcc) and mp64 == mp and np.array_equal(mean64, mean)
    pf = np.lib.stride_tricks.sliding_window_view(ref, 128)[::64]
    assert abs(mp - np.sqrt((pf * pf).sum(axis=1) / 128).max()) <= 1e-12
    assert np.allclose(mean, mfcc.mean(axis=0), rtol=1e-10, atol=1e-12)


def test_analyze_batch_equals_individual(ctx):
    """ss_sound_analyze_batch: many sounds back to back, framed on their own, one launch. Ragged lengths, odd offsets
    (misaligned 16-byte loads), sounds without a full MFCC frame or without a full power frame, an empty sound."""
    rng = np.random.default_rng(5)
    lens = [5000, 1023, 1024, 1025, 0, 127, 128, 44101, 3333, 256 * 40 + 1024 + 255]
    sounds = [rng.standard_normal(l) * 0.2 for l in lens]
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    mfcc, foff, mp, mean = ctx.analyze_batch(np.concatenate(sounds), off)
    assert int(foff[-1]) == mfcc.shape[0]
    for i, s in enumerate(sounds):
        m1, mp1, mean1 = ctx.analyze(s)
        assert int(foff[i + 1] - foff[i]) == m1.shape[0]
        assert np.array_equal(mfcc[int(foff[i]):int(foff[i + 1])], m1), i
        assert mp[i] == mp1, i
        if m1.shape[0]:
            assert np.allclose(mean[i], mean1, rtol=0, atol=1e-12), i
        else:
            assert np.all(np.isnan(mean[i])), i
    m0, f0, p0, a0 = ctx.analyze_batch(np.zeros(0), np.zeros(1, dtype=np.uint64))
    assert m0.shape == (0, 12) and f0.tolist() == [0] and p0.shape == (0,)


def test_dictionary_from_path_write_splits_and_timestamps(ctx, section71, tmp_path):
    """write_splits (src/lib.rs:155-178) -> SoundDictionary::from_path (src/sound.rs:304-320) round trip, and
    SoundSequence::from_timestamps (:419-430); both analyse all their sounds in one batch."""
    s16 = O.decode_pcm(section71["pcm"].astype(np.int32), 16)
    snd = api.Sound.from_samples(s16[:200000], 44100.0, None, "src", ctx)
    splits = [30000, 1500, 70000, 900, 97600]
    api.write_splits(snd, splits, str(tmp_path))
    (tmp_path / "notes.txt").write_text("not a wav")
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == ["00000_30000.wav", "00001_1500.wav", "00002_70000.wav", "00003_900.wav", "00004_97600.wav", "notes.txt"]
    d = api.SoundDictionary.from_path(str(tmp_path), ctx)
    assert [s.name for s in d.sounds] == [n[:-4] for n in names[:5]]
    pos = 0
    for s, ln in zip(d.sounds, splits):
        one = api.Sound.from_path(str(tmp_path / (s.name + ".wav")), ctx)
        assert len(s.samples()) == ln and np.array_equal(s.samples(), one.samples())
        assert np.array_equal(s.mfcc_arrays(), one.mfcc_arrays()) and s.max_power() == one.max_power()
        # 32-bit PCM of the writer: trunc(sample * i32::MAX) / i32::MAX
        assert np.array_equal(s.samples(), np.trunc(s16[pos:pos + ln] * 2147483647.0) / 2147483647.0)
        pos += ln
    stamps = [api.Timestamp(0.0, 0.5, "a"), api.Timestamp(0.5, 0.51, None), api.Timestamp(1.25, 2.0, "c")]
    seq = api.SoundSequence.from_timestamps(snd, stamps)
    assert [s.name for s in seq.sounds()] == ["a", None, "c"]
    for s, t in zip(seq.sounds(), stamps):
        a, b = int(round(t[0] * 44100)), int(round(t[1] * 44100))
        ref = api.Sound.from_samples(s16[a:b + 1], 44100.0, None, None, ctx)
        assert np.array_equal(s.samples(), ref.samples()) and np.array_equal(s.mfcc_arrays(), ref.mfcc_arrays())
    assert len(seq.distances()) == 2
    p = api.Partitioner.from_path(str(tmp_path / "00002_70000.wav"), ctx)
    assert p.depth == 5 and p.threshold == 4 and p.sound.num_frames() == (70000 - 1024) // 256 + 1
