"""ctypes binding of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs. The product package (soundsym_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "soundsym_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    L.orc_decode_pcm.argtypes = [_i32p, C.c_size_t, C.c_int, _f64p]
    L.orc_frame_count.restype = C.c_size_t
    L.orc_frame_count.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t]
    L.orc_hann.restype = C.c_double
    L.orc_hann.argtypes = [C.c_size_t, C.c_size_t]
    L.orc_mel_bins.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_size_t, _i32p]
    L.orc_dct.argtypes = [_f64p, C.c_int, _f64p]
    L.orc_mfcc.restype = C.c_size_t
    L.orc_mfcc.argtypes = [_f64p, C.c_size_t, C.c_double, C.c_int, C.c_size_t, C.c_size_t, C.c_double, C.c_double, C.c_void_p]
    L.orc_max_power.restype = C.c_double
    L.orc_max_power.argtypes = [_f64p, C.c_size_t]
    L.orc_mean_mfccs.argtypes = [_f64p, C.c_size_t, C.c_int, _f64p]
    L.orc_dot.restype = C.c_double
    L.orc_dot.argtypes = [_f64p, _f64p, C.c_size_t]
    L.orc_norm.restype = C.c_double
    L.orc_norm.argtypes = [_f64p, C.c_size_t]
    L.orc_cosine_sim.restype = C.c_double
    L.orc_cosine_sim.argtypes = [_f64p, C.c_size_t, _f64p, C.c_size_t]
    L.orc_cosine_sim_angular.restype = C.c_double
    L.orc_cosine_sim_angular.argtypes = [_f64p, _f64p, C.c_size_t]
    L.orc_cosine_match.argtypes = [_f64p, _u64p, C.c_size_t, _f64p, _u64p, C.c_size_t, C.c_void_p, _u32p, _f64p]
    L.orc_dtw.restype = C.c_double
    L.orc_dtw.argtypes = [_f64p, C.c_size_t, _f64p, C.c_size_t, C.c_int]
    L.orc_dtw_topk.argtypes = [_f64p, _u64p, C.c_size_t, _f64p, _u64p, C.c_size_t, C.c_int, C.c_int, _u32p, _f64p]
    L.orc_dtw_matrix.argtypes = [_f64p, _u64p, C.c_size_t, _f64p, _u64p, C.c_size_t, C.c_int, _f64p]
    L.orc_standardize.restype = C.c_int
    L.orc_standardize.argtypes = [_f64p, C.c_size_t, C.c_int, _f64p, _f64p, _f64p]
    L.orc_gmm_prepare.restype = C.c_int
    L.orc_gmm_prepare.argtypes = [_f64p, C.c_int, C.c_int, _f64p, _f64p]
    L.orc_gmm_posteriors.argtypes = [_f64p, C.c_size_t, C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p, _f64p]
    L.orc_max_index.restype = C.c_size_t
    L.orc_max_index.argtypes = [_f64p, C.c_size_t]
    L.orc_symbols.restype = C.c_int
    L.orc_symbols.argtypes = [_f64p, C.c_size_t, C.c_int, C.c_int, _f64p, _f64p, _f64p, _u8p]
    L.orc_gmm_train.restype = C.c_int
    L.orc_gmm_train.argtypes = [_f64p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, _f64p, _f64p, _f64p]
    L.orc_cast_votes.argtypes = [_u8p, C.c_size_t, C.c_int, _u32p]
    L.orc_split.restype = C.c_size_t
    L.orc_split.argtypes = [_u32p, C.c_size_t, C.c_int, _u64p]
    L.orc_resynth.argtypes = [_f64p, _u64p, _u32p, _u64p, C.c_size_t, _f64p]
    L.orc_num_threads.restype = C.c_int
    L.orc_hardware_threads.restype = C.c_int
    L.orc_set_threads.argtypes = [C.c_int]
    _LIB = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


# ---- constants of the reference, src/lib.rs:22-26 ------------------------------------------------------------------
NCOEFFS, NCLUSTERS, HOP, BIN = 12, 26, 256, 1024
F_LO, F_HI = 100.0, 8000.0


def set_threads(n):
    lib().orc_set_threads(int(n))


def hardware_threads():
    return int(lib().orc_hardware_threads())


def decode_pcm(pcm, bits):
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    out = np.empty(pcm.shape[0], dtype=np.float64)
    lib().orc_decode_pcm(pcm, pcm.shape[0], int(bits), out)
    return out


def frame_count(n, bin=BIN, hop=HOP):
    return int(lib().orc_frame_count(n, bin, hop))


def mel_bins(ncoeffs=NCOEFFS, sr=44100.0, bin=BIN, f_lo=F_LO, f_hi=F_HI):
    out = np.empty(ncoeffs + 2, dtype=np.int32)
    lib().orc_mel_bins(ncoeffs, f_lo, f_hi, sr, bin, out)
    return out


def dct(e):
    e = _f64(e)
    out = np.empty_like(e)
    lib().orc_dct(e, e.shape[0], out)
    return out


def mfcc(samples, sr=44100.0, ncoeffs=NCOEFFS, bin=BIN, hop=HOP, f_lo=F_LO, f_hi=F_HI):
    samples = _f64(samples)
    frames = frame_count(samples.shape[0], bin, hop)
    out = np.empty((frames, ncoeffs), dtype=np.float64)
    if frames:
        lib().orc_mfcc(samples, samples.shape[0], sr, ncoeffs, bin, hop, f_lo, f_hi, out.ctypes.data_as(C.c_void_p))
    return out


def max_power(samples):
    samples = _f64(samples)
    return float(lib().orc_max_power(samples, samples.shape[0]))


def mean_mfccs(m):
    m = _f64(m)
    out = np.empty(m.shape[1], dtype=np.float64)
    with np.errstate(all="ignore"):
        lib().orc_mean_mfccs(m, m.shape[0], m.shape[1], out)
    return out


def cosine_sim(me, you):
    me, you = _f64(me).ravel(), _f64(you).ravel()
    return float(lib().orc_cosine_sim(me, me.shape[0], you, you.shape[0]))


def cosine_sim_angular(me, you):
    me, you = _f64(me).ravel(), _f64(you).ravel()
    return float(lib().orc_cosine_sim_angular(me, you, me.shape[0]))


def cosine_match(dict_flat, dict_off_frames, q_flat, q_off_frames, c, targets=None):
    """reference matcher (src/sound.rs:351-370). Offsets are in FRAMES; returns (idx u32[nq], dist f64[nq])."""
    d, q = _f64(dict_flat).ravel(), _f64(q_flat).ravel()
    do, qo = _u64(np.asarray(dict_off_frames) * c), _u64(np.asarray(q_off_frames) * c)
    nq = qo.shape[0] - 1
    idx = np.empty(nq, dtype=np.uint32)
    dist = np.empty(nq, dtype=np.float64)
    t = None
    if targets is not None:
        t = _f64(targets)
    lib().orc_cosine_match(d, do, do.shape[0] - 1, q, qo, nq, t.ctypes.data_as(C.c_void_p) if t is not None else None, idx, dist)
    return idx, dist


def dtw(a, b):
    a, b = _f64(a), _f64(b)
    c = a.shape[1] if a.ndim == 2 else b.shape[1]
    return float(lib().orc_dtw(a, a.shape[0], b, b.shape[0], c))


def dtw_topk(dict_flat, dict_off_frames, q_flat, q_off_frames, c, k=1):
    d, q = _f64(dict_flat).ravel(), _f64(q_flat).ravel()
    do, qo = _u64(dict_off_frames), _u64(q_off_frames)
    nq = qo.shape[0] - 1
    idx = np.empty((nq, k), dtype=np.uint32)
    dist = np.empty((nq, k), dtype=np.float64)
    lib().orc_dtw_topk(d, do, do.shape[0] - 1, q, qo, nq, c, k, idx, dist)
    return idx, dist


def dtw_matrix(dict_flat, dict_off_frames, q_flat, q_off_frames, c):
    """every (query, segment) DTW distance, f64 [nq, nseg]"""
    d, q = _f64(dict_flat).ravel(), _f64(q_flat).ravel()
    do, qo = _u64(dict_off_frames), _u64(q_off_frames)
    out = np.empty((qo.shape[0] - 1, do.shape[0] - 1), dtype=np.float64)
    lib().orc_dtw_matrix(d, do, do.shape[0] - 1, q, qo, qo.shape[0] - 1, c, out)
    return out


def standardize(x):
    x = _f64(x)
    out = np.empty_like(x)
    mean = np.empty(x.shape[1])
    std = np.empty(x.shape[1])
    rc = lib().orc_standardize(x, x.shape[0], x.shape[1], out, mean, std)
    if rc != 0:
        raise ValueError("standardize needs >= 2 rows")
    return out, mean, std


def gmm_posteriors(z, means, covs, weights):
    z, means, covs, weights = _f64(z), _f64(means), _f64(covs), _f64(weights)
    ncomp, c = means.shape
    inv = np.empty_like(covs)
    sd = np.empty(ncomp)
    rc = lib().orc_gmm_prepare(covs, ncomp, c, inv, sd)
    if rc != 0:
        raise ValueError("singular covariance in component %d" % (-rc - 1))
    out = np.empty((z.shape[0], ncomp))
    with np.errstate(all="ignore"):
        lib().orc_gmm_posteriors(z, z.shape[0], c, ncomp, means, inv, sd, weights, out)
    return out


def symbols(mfcc_rows, means, covs, weights):
    m, means, covs, weights = _f64(mfcc_rows), _f64(means), _f64(covs), _f64(weights)
    out = np.empty(m.shape[0], dtype=np.uint8)
    rc = lib().orc_symbols(m, m.shape[0], m.shape[1], means.shape[0], means, covs, weights, out)
    if rc != 0:
        raise ValueError("orc_symbols failed rc=%d" % rc)
    return out


def gmm_train(z, ncomp=NCLUSTERS, iters=5, reg=0.1, seed=0):
    """seeded restatement of train_model's EM (src/lib.rs:44-54) on ALREADY standardised rows; retries like the
    reference's `while let Err` loop with seed+1, seed+2, ..."""
    z = _f64(z)
    c = z.shape[1]
    means = np.empty((ncomp, c))
    covs = np.empty((ncomp, c, c))
    weights = np.empty(ncomp)
    for attempt in range(64):
        rc = lib().orc_gmm_train(z, z.shape[0], c, ncomp, iters, reg, seed + attempt, means, covs, weights)
        if rc == 0:
            return means, covs, weights
    raise RuntimeError("gmm_train failed rc=%d" % rc)


def cast_votes(sym, depth):
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    votes = np.zeros(sym.shape[0] + 1, dtype=np.uint32)
    lib().orc_cast_votes(sym, sym.shape[0], depth, votes)
    return votes


def split(votes, n, threshold):
    votes = np.ascontiguousarray(votes, dtype=np.uint32)
    out = np.empty(max(n, 1), dtype=np.uint64)
    nseg = lib().orc_split(votes, n, threshold, out)
    return out[:nseg].copy()


def partition(mfcc_rows, model, depth, threshold, hop=HOP):
    """Partitioner::partition_other (src/lib.rs:112-144): segment lengths in samples."""
    sym = symbols(mfcc_rows, *model)
    votes = cast_votes(sym, depth)
    return split(votes, sym.shape[0], threshold) * np.uint64(hop), sym, votes


def resynth(dict_samples, dict_off, match_idx, tgt_len):
    ds, do = _f64(dict_samples), _u64(dict_off)
    mi = np.ascontiguousarray(match_idx, dtype=np.uint32)
    tl = _u64(tgt_len)
    out = np.empty(int(tl.sum()), dtype=np.float64)
    lib().orc_resynth(ds, do, mi, tl, mi.shape[0], out)
    return out
