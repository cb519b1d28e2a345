"""The C++ host mirror (include/soundsym.hpp) driven like the reference's examples/reconstruction.rs and partition.rs,
compared with the golden fixtures. GPU test: the binary calls the CUDA library through the C ABI."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "reconstruction")


def write_wav(path, pcm, bits, sr=44100):
    pcm = np.asarray(pcm)
    if bits == 16:
        body = pcm.astype("<i2").tobytes()
    elif bits == 24:
        v = pcm.astype(np.int32) & 0xFFFFFF
        body = np.stack([v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8).tobytes()
    else:
        raise ValueError(bits)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * bits // 8, bits // 8, bits)
                + b"data" + struct.pack("<I", len(body)) + body)


def write_model(path, g):
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", g["gmm_means"].shape[0], g["gmm_means"].shape[1]))
        for k in ("gmm_means", "gmm_covs", "gmm_weights"):
            f.write(np.ascontiguousarray(g[k], dtype="<f8").tobytes())


def run(args):
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    p = subprocess.run([BIN] + args, capture_output=True, text=True, timeout=300)
    return p.returncode, json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
def test_cpp_partition_and_reconstruction(tmp_path, section71, sample_excerpt):
    src, tgt, mdl, out = (str(tmp_path / n) for n in ("source.wav", "target.wav", "model.bin", "out.wav"))
    write_wav(src, section71["pcm"], 16)
    write_wav(tgt, sample_excerpt["pcm"], 24)
    write_model(mdl, section71)
    # examples/partition.rs defaults: threshold 3, depth 4
    rc, r = run(["-s", src, "-m", mdl, "--depth", "4", "--threshold", "3", "--partition-only"])
    assert rc == 0 and r["source_frames"] == 1978 and r["max_power"] == float(section71["max_power"])
    assert r["splits"] == [int(x) for x in section71["splits_d4t3"]]
    # examples/reconstruction.rs: threshold 4, depth 3
    rc, r = run(["-s", src, "-t", tgt, "-o", out, "-m", mdl])
    assert rc == 0 and r["nsplits"] == len(section71["splits_d3t4"]) and r["ntarget"] == len(sample_excerpt["splits_d3t4"])
    assert r["idx"] == [int(x) for x in sample_excerpt["cos_idx"]]
    assert r["out_samples"] == int(sample_excerpt["resynth_len"]) and abs(r["out_sum"] - float(sample_excerpt["resynth_sum"])) < 1e-9
    assert os.path.getsize(out) == 44 + 4 * r["out_samples"]
    rc, r = run(["-s", src, "-t", tgt, "-m", mdl, "--dtw"])
    assert rc == 0 and r["idx"] == [int(x) for x in sample_excerpt["dtw_idx"][:, 0]]
    # Partitioner without a model: CosError("Must first train model"), src/lib.rs:140-142
    rc, r = run(["-s", src, "--partition-only"])
    assert rc == 1 and r["code"] == -4 and "Must first train model" in r["error"]


def test_cpp_host_mirror_compiles():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-B"], stdout=subprocess.DEVNULL)
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_cpp_matcher_flow_from_directories(tmp_path, section71):
    """examples/matcher.rs through the C++ mirror: write_splits (examples/partition.rs) -> SoundDictionary::from_path on the
    directory of cuts (batched analysis) -> every query file matched in one call; compared with the Python mirror, which
    shares nothing with it above the C ABI."""
    from oracle import oracle as O
    from soundsym_b200 import api
    src, mdl = str(tmp_path / "source.wav"), str(tmp_path / "model.bin")
    ddir, qdir = tmp_path / "dict", tmp_path / "queries"
    ddir.mkdir(), qdir.mkdir()
    write_wav(src, section71["pcm"], 16)
    write_model(mdl, section71)
    rc, r = run(["-s", src, "-m", mdl, "--depth", "4", "--threshold", "3", "--partition-only", "--write-splits", str(ddir)])
    splits = [int(x) for x in section71["splits_d4t3"]]
    assert rc == 0 and r["splits"] == splits
    files = sorted(os.listdir(ddir))
    assert files == ["%05d_%d.wav" % (i, s) for i, s in enumerate(splits)]
    # queries: a few cuts of the same recording at shifted positions, one of them silent
    s16 = O.decode_pcm(section71["pcm"].astype(np.int32), 16)
    rng = np.random.default_rng(3)
    for i in range(6):
        a = int(rng.integers(0, len(s16) - 30000))
        write_wav(str(qdir / ("q%d.wav" % i)), section71["pcm"][a:a + int(rng.integers(3000, 20000))], 16)
    write_wav(str(qdir / "silence.wav"), np.zeros(5000, dtype=np.int16), 16)
    rc, r = run(["--matcher", str(ddir), str(qdir)])
    ctx = api.Context(0)
    d = api.SoundDictionary.from_path(str(ddir), ctx)
    q = api.SoundDictionary.from_path(str(qdir), ctx)
    loud = [s for s in q.sounds if not s.max_power() < 0.03]  # examples/matcher.rs:40
    assert 1 <= len(loud) < 7  # the all-zero file is silent, the loudest passages are not
    assert rc == 0 and r["dictionary"] == len(splits) and r["queries"] == 7 and r["silent"] == 7 - len(loud)
    idx, _ = d.match_indices(loud)
    assert r["matches"] == [[s.name, d.sounds[int(i)].name] for s, i in zip(loud, idx[:, 0])]
