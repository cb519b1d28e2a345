// oracle/soundsym_oracle.cpp — CPU restatement (f64; single thread unless orc_set_threads(n>1)) of the soundsym hot path.
//
// TEST INFRASTRUCTURE ONLY: the product (soundsym_b200/csrc) never links, loads or calls this file. It is loaded by
// tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / `--impl reference` legs, always as the checker
// or the timed CPU baseline, never as a fallback.
//
// The reference cannot be built here (Rust; five un-vendored crates, two un-pinned), so this file restates its
// loops in the same fold orders and in f64. Every function cites the reference lines it follows
// (paths relative to /root/reference). Behaviour that lives in absent third-party crates is fixed by the flags in
// ASSUMPTIONS.h (A1..A9) and marked [RECALL]. DTW (A8) has no reference counterpart: PARITY UNPINNED.
//
// Build: g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -pthread (see oracle/Makefile). -ffp-contract=off keeps
// mul and add as separate IEEE operations, as rustc does.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>
#include <atomic>
#include <thread>

#include "ASSUMPTIONS.h"

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

// ---------------------------------------------------------------------------------------------------------------
// FFT: iterative radix-2, f64. vox_box calls rustfft on the windowed frame (full complex transform) [RECALL];
// any exact-arithmetic-equivalent DFT differs from it only in rounding (~1e-16 relative).
void fft_inplace(std::vector<std::complex<double>>& x) {
    const size_t n = x.size();
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(x[i], x[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < len / 2; k++) {
                const double ang = -2.0 * kPi * (double)k / (double)len;
                const std::complex<double> w(std::cos(ang), std::sin(ang));
                const std::complex<double> u = x[i + k], v = x[i + k + len / 2] * w;
                x[i + k] = u + v;
                x[i + k + len / 2] = u - v;
            }
        }
    }
}

// multi-thread helper for the timed CPU-baseline entries (the reference itself is single-threaded; orc_set_threads(1)
// reproduces that). Dynamic chunking over [0, n).
int g_threads = 1;
template <typename F>
void parallel_for(size_t n, size_t chunk, F body) {
    const int nt = g_threads;
    if (nt <= 1 || n <= chunk) {
        for (size_t i = 0; i < n; i++) body(i);
        return;
    }
    std::atomic<size_t> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&] {
            for (;;) {
                const size_t b = next.fetch_add(chunk);
                if (b >= n) return;
                const size_t e = b + chunk < n ? b + chunk : n;
                for (size_t i = b; i < e; i++) body(i);
            }
        });
    for (auto& th : pool) th.join();
}

inline double hz_to_mel(double hz) { return 1125.0 * std::log1p(hz / 700.0); }  // [RECALL] vox_box::spectrum
inline double mel_to_hz(double mel) { return 700.0 * (std::exp(mel / 1125.0) - 1.0); }

}  // namespace

extern "C" {

// ---- row 1: PCM decode, src/sound.rs:117-120 --------------------------------------------------------------------
// s as f64 / (i32::MAX >> (32 - bits)) as f64
void orc_decode_pcm(const int32_t* pcm, size_t n, int bits, double* out) {
    const double denom = (double)(INT32_MAX >> (32 - bits));
    for (size_t i = 0; i < n; i++) out[i] = (double)pcm[i] / denom;
}

// ---- row 4: framing rule of sample::window::Windower (A1), call sites src/sound.rs:228-229, 245-246 -------------
size_t orc_frame_count(size_t n, size_t bin, size_t hop) { return n >= bin ? (n - bin) / hop + 1 : 0; }

// Hann window value (A2), sample::window::Hanning [RECALL]
double orc_hann(size_t i, size_t bin) {
#if ORC_A2_HANN_SYMMETRIC
    return 0.5 * (1.0 - std::cos(2.0 * kPi * (double)i / (double)(bin - 1)));
#else
    return 0.5 * (1.0 - std::cos(2.0 * kPi * (double)i / (double)bin));
#endif
}

// mel band edges as FFT bin indices (A3b-c): ncoeffs+2 entries
void orc_mel_bins(int ncoeffs, double f_lo, double f_hi, double sample_rate, size_t bin, int* out_bins) {
    const double lo = hz_to_mel(f_lo), range = hz_to_mel(f_hi) - lo;
    for (int i = 0; i < ncoeffs + 2; i++) {
        const double point = ((double)i / (double)ncoeffs) * range + lo;
        out_bins[i] = (int)std::floor((double)(bin + 1) * mel_to_hz(point) / sample_rate);
    }
}

// DCT-II x2 (A3g): c_k = 2 * sum_n e_n cos(pi k (2n+1) / (2C)). vox_box KAT: dct([.2,.3,.4,.3]) = [2.4,-0.26131,-0.28284,0.10823]
void orc_dct(const double* e, int c, double* out) {
    for (int k = 0; k < c; k++) {
        double acc = 0.0;
        for (int n = 0; n < c; n++) acc = acc + e[n] * std::cos(kPi * (double)k * (2.0 * (double)n + 1.0) / (2.0 * (double)c));
        out[k] = 2.0 * acc;
    }
}

// ---- rows 3-5: analyze_mfccs, src/sound.rs:215-242 ----------------------------------------------------------------
// frames of `bin` every `hop`, Hann-weighted (A2), each -> ncoeffs MFCCs over (f_lo, f_hi) (A3, A4); row-major frames x C.
// Returns the number of frames. out may be NULL to query the count.
size_t orc_mfcc(const double* samples, size_t n, double sample_rate, int ncoeffs, size_t bin, size_t hop, double f_lo,
                double f_hi, double* out) {
    const size_t frames = orc_frame_count(n, bin, hop);
    if (!out || frames == 0) return frames;
    std::vector<int> bins(ncoeffs + 2);
    orc_mel_bins(ncoeffs, f_lo, f_hi, sample_rate, bin, bins.data());
    std::vector<double> win(bin);
    for (size_t i = 0; i < bin; i++) win[i] = orc_hann(i, bin);
    std::vector<std::complex<double>> buf(bin);
    std::vector<double> energies(ncoeffs);
    for (size_t f = 0; f < frames; f++) {
        const double* s = samples + f * hop;
        for (size_t i = 0; i < bin; i++) buf[i] = std::complex<double>(s[i] * win[i], 0.0);
        fft_inplace(buf);
        for (int b = 0; b < ncoeffs; b++) {
            const int b0 = bins[b], b1 = bins[b + 1], b2 = bins[b + 2];
            const int up = b1 - b0, down = b2 - b1;
            double up_sum = 0.0, down_sum = 0.0;
            for (int k = b0, i = 0; k < b1; k++, i++) {
                const size_t kk = (size_t)k % bin;
#if ORC_A3_POWER_SPECTRUM
                const double p = std::norm(buf[kk]);
#else
                const double p = std::abs(buf[kk]);
#endif
                up_sum = up_sum + p * ((double)i / (double)up);
            }
            for (int k = b1, i = 0; k < b2; k++, i++) {
                const size_t kk = (size_t)k % bin;
#if ORC_A3_POWER_SPECTRUM
                const double p = std::norm(buf[kk]);
#else
                const double p = std::abs(buf[kk]);
#endif
                down_sum = down_sum + p * (1.0 - (double)i / (double)down);
            }
            const double e = up_sum + down_sum;
            energies[b] = std::log10(e > ORC_A4_ENERGY_FLOOR ? e : ORC_A4_ENERGY_FLOOR);
        }
        orc_dct(energies.data(), ncoeffs, out + f * ncoeffs);
    }
    return frames;
}

// ---- row 6: analyze_max_power, src/sound.rs:244-256 --------------------------------------------------------------
// rectangular frames 128/64; per frame sqrt(sum s^2 / count); max over frames starting from 0.0
double orc_max_power(const double* samples, size_t n) {
    const size_t frames = orc_frame_count(n, 128, 64);
    double best = 0.0;
    for (size_t f = 0; f < frames; f++) {
        double acc = 0.0;
        size_t count = 0;
        for (size_t i = 0; i < 128; i++) {
            const double s = samples[f * 64 + i];
            count += 1;
            acc = acc + s * s;
        }
        const double rms = std::sqrt(acc / (double)count);
        best = std::fmax(best, rms);  // f64::max: NaN-ignoring like fmax
    }
    return best;
}

// ---- row 7: analyze_mean_mfccs, src/sound.rs:271-286 --------------------------------------------------------------
void orc_mean_mfccs(const double* mfcc, size_t frames, int c, double* out) {
    std::vector<double> sums(c, 0.0);
    for (size_t f = 0; f < frames; f++)
        for (int k = 0; k < c; k++) sums[k] += mfcc[f * c + k];
    for (int k = 0; k < c; k++) out[k] = sums[k] / (double)frames;  // 0/0 = NaN for an empty sound, as the reference
}

// ---- row 15: rulinalg::utils::dot (A9) [RECALL] -------------------------------------------------------------------
double orc_dot(const double* xs, const double* ys, size_t len) {
#if ORC_A9_DOT_UNROLL8
    double p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0, s = 0;
    while (len >= 8) {
        p0 = p0 + xs[0] * ys[0];
        p1 = p1 + xs[1] * ys[1];
        p2 = p2 + xs[2] * ys[2];
        p3 = p3 + xs[3] * ys[3];
        p4 = p4 + xs[4] * ys[4];
        p5 = p5 + xs[5] * ys[5];
        p6 = p6 + xs[6] * ys[6];
        p7 = p7 + xs[7] * ys[7];
        xs += 8;
        ys += 8;
        len -= 8;
    }
    s = s + (p0 + p4);
    s = s + (p1 + p5);
    s = s + (p2 + p6);
    s = s + (p3 + p7);
    for (size_t i = 0; i < len; i++) s = s + xs[i] * ys[i];
    return s;
#else
    double s = 0;
    for (size_t i = 0; i < len; i++) s = s + xs[i] * ys[i];
    return s;
#endif
}

// norm, src/sound.rs:36-38: sequential left fold item*item + memo, NO sqrt
double orc_norm(const double* me, size_t len) {
    double memo = 0.0;
    for (size_t i = 0; i < len; i++) memo = me[i] * me[i] + memo;
    return memo;
}

// cosine_sim, src/sound.rs:23-33
double orc_cosine_sim(const double* me, size_t me_len, const double* you, size_t you_len) {
    const size_t len = me_len < you_len ? me_len : you_len;
    const double nrm = orc_norm(me, me_len) * orc_norm(you, you_len);
    const double dot = orc_dot(me, you, len);
    return dot / nrm;
}

// cosine_sim_angular, src/sound.rs:62-69 (x > 1 -> 1, x < -1 -> 1 (sic), acos(x)/pi)
double orc_cosine_sim_angular(const double* me, const double* you, size_t c) {
    double sim = orc_cosine_sim(me, c, you, c);
    if (sim > 1.0) sim = 1.0;
    else if (sim < -1.0) sim = 1.0;
    return std::acos(sim) * 0.318309886183790671537767526745028724;  // f64::consts::FRAC_1_PI
}

// ---- row 16: SoundDictionary::at_distance, src/sound.rs:351-370 -----------------------------------------------------
// dict = ragged list (flat values + value offsets, nseg+1 entries, in f64 VALUES not frames);
// returns argmin_d |sim_d - target| with strict '<' from (0, 2.0); out_dist = that minimum (2.0 if nothing won).
size_t orc_at_distance(const double* dict, const uint64_t* dict_off, size_t nseg, const double* q, size_t q_len,
                       double target, double* out_dist) {
    size_t min_idx = 0;
    double min_distance = 2.0;
    for (size_t d = 0; d < nseg; d++) {
        const double sim = orc_cosine_sim(dict + dict_off[d], (size_t)(dict_off[d + 1] - dict_off[d]), q, q_len);
        const double distance = std::fabs(sim - target);
        if (distance < min_distance) {
            min_idx = d;
            min_distance = distance;
        }
    }
    if (out_dist) *out_dist = min_distance;
    return min_idx;
}

// batched form: nq queries (ragged), one target each (targets may be NULL => 1.0 = match_sound, src/sound.rs:346-348)
void orc_cosine_match(const double* dict, const uint64_t* dict_off, size_t nseg, const double* q, const uint64_t* q_off,
                      size_t nq, const double* targets, uint32_t* out_idx, double* out_dist) {
    parallel_for(nq, 4, [&](size_t i) {
        double dist;
        out_idx[i] = (uint32_t)orc_at_distance(dict, dict_off, nseg, q + q_off[i], (size_t)(q_off[i + 1] - q_off[i]),
                                               targets ? targets[i] : 1.0, &dist);
        if (out_dist) out_dist[i] = dist;
    });
}

// ---- row 20: DTW (A8). frames x c row-major; returns D(Lq-1,Ld-1)/(Lq+Ld); +inf if either side is empty ------------
double orc_dtw(const double* a, size_t la, const double* b, size_t lb, int c) {
    if (la == 0 || lb == 0) return std::numeric_limits<double>::infinity();
    static thread_local std::vector<double> prev, cur;  // row buffers reused across calls (same values, no per-pair allocation)
    if (prev.size() < lb) prev.resize(lb), cur.resize(lb);
    for (size_t i = 0; i < la; i++) {
        for (size_t j = 0; j < lb; j++) {
            double cost = 0.0;
            for (int k = 0; k < c; k++) {
                const double d = a[i * c + k] - b[j * c + k];
                cost = cost + d * d;
            }
            double m;
            if (i == 0 && j == 0) m = 0.0;
            else if (i == 0) m = cur[j - 1];
            else if (j == 0) m = prev[0];
            else m = std::fmin(std::fmin(prev[j], cur[j - 1]), prev[j - 1]);
            cur[j] = cost + m;
        }
        std::swap(prev, cur);
    }
    return prev[lb - 1] / (double)(la + lb);
}

// every (query, segment) distance: out[nq x nseg]. Checker for the scan-error measurement (tests/test_parity_large_gpu.py).
void orc_dtw_matrix(const double* dict, const uint64_t* dict_off, size_t nseg, const double* q, const uint64_t* q_off, size_t nq, int c,
                    double* out) {
    parallel_for(nq, 1, [&](size_t qi) {
        const double* qa = q + q_off[qi] * c;
        const size_t lq = (size_t)(q_off[qi + 1] - q_off[qi]);
        for (size_t d = 0; d < nseg; d++)
            out[qi * nseg + d] = orc_dtw(qa, lq, dict + dict_off[d] * c, (size_t)(dict_off[d + 1] - dict_off[d]), c);
    });
}

// top-k DTW match: frame offsets (nseg+1 entries, in FRAMES). For each query the k smallest (distance, index)
// lexicographic, ascending; NaN never wins; unfilled slots = (inf, 0xFFFFFFFF).
void orc_dtw_topk(const double* dict, const uint64_t* dict_off, size_t nseg, const double* q, const uint64_t* q_off,
                  size_t nq, int c, int k, uint32_t* out_idx, double* out_dist) {
    parallel_for(nq, 1, [&](size_t qi) {
        std::vector<std::pair<double, uint32_t>> best;  // sorted ascending, size <= k
        const double* qa = q + q_off[qi] * c;
        const size_t lq = (size_t)(q_off[qi + 1] - q_off[qi]);
        for (size_t d = 0; d < nseg; d++) {
            const double dist = orc_dtw(qa, lq, dict + dict_off[d] * c, (size_t)(dict_off[d + 1] - dict_off[d]), c);
            if (!(dist < std::numeric_limits<double>::infinity())) continue;  // NaN / inf never win
            std::pair<double, uint32_t> cand(dist, (uint32_t)d);
            if ((int)best.size() < k) {
                best.insert(std::upper_bound(best.begin(), best.end(), cand), cand);
            } else if (cand < best.back()) {
                best.pop_back();
                best.insert(std::upper_bound(best.begin(), best.end(), cand), cand);
            }
        }
        for (int s = 0; s < k; s++) {
            out_idx[qi * k + s] = s < (int)best.size() ? best[s].second : 0xFFFFFFFFu;
            out_dist[qi * k + s] = s < (int)best.size() ? best[s].first : std::numeric_limits<double>::infinity();
        }
    });
}

// total DP cells of a ragged all-pairs match (the unit of the headline metric)
uint64_t orc_dtw_cells(const uint64_t* dict_off, size_t nseg, const uint64_t* q_off, size_t nq) {
    const uint64_t dsum = dict_off[nseg] - dict_off[0], qsum = q_off[nq] - q_off[0];
    return dsum * qsum;
}

// ---- rows 9-10: Standardizer (A5) and GMM (A6), src/lib.rs:44-60 ---------------------------------------------------
// fit on the data, per-column mean and (n - ddof) variance; out = (x - mean)/sqrt(var). Returns 0, or -1 if rows < 2.
int orc_standardize(const double* x, size_t rows, int c, double* out, double* out_mean, double* out_std) {
    if (rows < 2) return -1;
    std::vector<double> mean(c, 0.0), var(c, 0.0);
    for (size_t r = 0; r < rows; r++)
        for (int k = 0; k < c; k++) mean[k] += x[r * c + k];
    for (int k = 0; k < c; k++) mean[k] /= (double)rows;
    for (size_t r = 0; r < rows; r++)
        for (int k = 0; k < c; k++) {
            const double d = mean[k] - x[r * c + k];
            var[k] += d * d;
        }
    for (int k = 0; k < c; k++) var[k] /= (double)(rows - ORC_A5_VARIANCE_DDOF);
    for (int k = 0; k < c; k++) {
        const double sd = std::sqrt(var[k]);
        if (out_mean) out_mean[k] = mean[k];
        if (out_std) out_std[k] = sd;
        if (out)
            for (size_t r = 0; r < rows; r++) out[r * c + k] = (x[r * c + k] - mean[k]) / sd;
    }
    return 0;
}

}  // extern "C"

namespace {

// LU with partial pivoting (rulinalg PartialPivLu [RECALL]): inverse and determinant of a c x c matrix (row-major).
bool lu_inv_det(const double* a, int c, double* inv, double* det) {
    std::vector<double> lu(a, a + c * c);
    std::vector<int> perm(c);
    for (int i = 0; i < c; i++) perm[i] = i;
    double d = 1.0;
    for (int col = 0; col < c; col++) {
        int piv = col;
        double best = std::fabs(lu[col * c + col]);
        for (int r = col + 1; r < c; r++)
            if (std::fabs(lu[r * c + col]) > best) best = std::fabs(lu[r * c + col]), piv = r;
        if (best == 0.0) return false;
        if (piv != col) {
            for (int k = 0; k < c; k++) std::swap(lu[piv * c + k], lu[col * c + k]);
            std::swap(perm[piv], perm[col]);
            d = -d;
        }
        d *= lu[col * c + col];
        for (int r = col + 1; r < c; r++) {
            lu[r * c + col] /= lu[col * c + col];
            const double f = lu[r * c + col];
            for (int k = col + 1; k < c; k++) lu[r * c + k] -= f * lu[col * c + k];
        }
    }
    *det = d;
    for (int e = 0; e < c; e++) {  // solve A x = unit_e
        std::vector<double> y(c);
        for (int i = 0; i < c; i++) {
            double s = perm[i] == e ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) s -= lu[i * c + k] * y[k];
            y[i] = s;
        }
        for (int i = c - 1; i >= 0; i--) {
            double s = y[i];
            for (int k = i + 1; k < c; k++) s -= lu[i * c + k] * inv[k * c + e];
            inv[i * c + e] = s / lu[i * c + i];
        }
    }
    return true;
}

}  // namespace

extern "C" {

// precision (inverse covariance) matrices and sqrt(det) per component; returns 0 or -(j+1) if component j is singular
int orc_gmm_prepare(const double* covs, int ncomp, int c, double* out_inv, double* out_sqrt_det) {
    for (int j = 0; j < ncomp; j++) {
        double det;
        if (!lu_inv_det(covs + (size_t)j * c * c, c, out_inv + (size_t)j * c * c, &det)) return -(j + 1);
        out_sqrt_det[j] = std::sqrt(det);
    }
    return 0;
}

// GaussianMixtureModel::predict on ALREADY standardised rows (A6) -> rows x ncomp posteriors
// pdf_ij = exp(-0.5 * (x-mu)^T P (x-mu)) / sqrt_det_j, evaluated as (diff * P) * diff^T like the matrix chain
// `&diff * &cov_inv * diff.transpose()`; weighted sum via dot(mix_weights, pdfs) (A9).
void orc_gmm_posteriors(const double* z, size_t rows, int c, int ncomp, const double* means, const double* inv,
                        const double* sqrt_det, const double* weights, double* out) {
    std::vector<double> pdfs(ncomp), tmp(c);
    for (size_t r = 0; r < rows; r++) {
        for (int j = 0; j < ncomp; j++) {
            const double* P = inv + (size_t)j * c * c;
            double quad = 0.0;
            for (int col = 0; col < c; col++) {  // tmp = diff (1 x c) * P (c x c)
                double s = 0.0;
                for (int k = 0; k < c; k++) s = s + (z[r * c + k] - means[j * c + k]) * P[k * c + col];
                tmp[col] = s;
            }
            for (int col = 0; col < c; col++) quad = quad + tmp[col] * (z[r * c + col] - means[j * c + col]);
            pdfs[j] = std::exp(quad * -0.5) / sqrt_det[j];
        }
        const double wsum = orc_dot(weights, pdfs.data(), (size_t)ncomp);
        for (int j = 0; j < ncomp; j++) out[r * ncomp + j] = weights[j] * pdfs[j] / wsum;
    }
}

// max_index, src/sound.rs:486-495: first index whose value is > the running max, starting from (0, 0.0)
size_t orc_max_index(const double* vals, size_t n) {
    size_t max_idx = 0;
    double max_dist = 0.0;
    for (size_t i = 0; i < n; i++)
        if (vals[i] > max_dist) max_idx = i, max_dist = vals[i];
    return max_idx;
}

// rows 10-11 together: discretize_with_model + symbolisation loop, src/lib.rs:56-60, 123-131.
// raw MFCC rows -> standardise with THIS input's stats -> posteriors -> 'A' + max_index. Returns 0 / <0 on error.
int orc_symbols(const double* mfcc, size_t rows, int c, int ncomp, const double* means, const double* covs,
                const double* weights, uint8_t* out_sym) {
    std::vector<double> z(rows * c), inv((size_t)ncomp * c * c), sd(ncomp), post(rows * ncomp);
    if (orc_standardize(mfcc, rows, c, z.data(), nullptr, nullptr) != 0) return -1;
    const int rc = orc_gmm_prepare(covs, ncomp, c, inv.data(), sd.data());
    if (rc != 0) return rc - 100;
    orc_gmm_posteriors(z.data(), rows, c, ncomp, means, inv.data(), sd.data(), weights, post.data());
    for (size_t r = 0; r < rows; r++) out_sym[r] = (uint8_t)('A' + orc_max_index(post.data() + r * ncomp, ncomp));
    return 0;
}

// row 9: train_model, src/lib.rs:44-54 — EM for `ncomp` full-covariance Gaussians on standardised data,
// CovOption::Regularized(reg) [RECALL rusty-machine 0.5.4 gmm.rs]: initial covariances = data covariance/(n-1) + reg*I
// for every component, means = `ncomp` distinct random rows (the reference uses thread_rng -> non-deterministic; here a
// seeded mt19937_64 so a model can be reproduced), `iters` rounds of membership_weights + update_params, where the new
// covariance is (sum_i w_ik diff diff^T + reg*I) / sum_i w_ik. Returns 0, or <0 if a covariance became singular.
int orc_gmm_train(const double* z, size_t rows, int c, int ncomp, int iters, double reg, uint64_t seed, double* means,
                  double* covs, double* weights) {
    if (rows < (size_t)ncomp || rows < 2) return -1;
    std::vector<double> colmean(c, 0.0);
    for (size_t r = 0; r < rows; r++)
        for (int k = 0; k < c; k++) colmean[k] += z[r * c + k];
    for (int k = 0; k < c; k++) colmean[k] /= (double)rows;
    std::vector<double> cov0((size_t)c * c, 0.0);
    for (int j = 0; j < c; j++)
        for (int k = 0; k < c; k++) {
            double s = 0.0;
            for (size_t r = 0; r < rows; r++) s += (z[r * c + j] - colmean[j]) * (z[r * c + k] - colmean[k]);
            cov0[j * c + k] = s * (1.0 / (double)(rows - 1));
        }
    for (int k = 0; k < c; k++) cov0[k * c + k] += reg;
    for (int j = 0; j < ncomp; j++) std::memcpy(covs + (size_t)j * c * c, cov0.data(), sizeof(double) * c * c);
    std::mt19937_64 rng(seed);
    std::vector<size_t> idx(rows);
    for (size_t i = 0; i < rows; i++) idx[i] = i;
    for (int j = 0; j < ncomp; j++) {  // partial Fisher-Yates: ncomp distinct rows
        std::uniform_int_distribution<size_t> pick(j, rows - 1);
        std::swap(idx[j], idx[pick(rng)]);
        std::memcpy(means + (size_t)j * c, z + idx[j] * c, sizeof(double) * c);
    }
    for (int j = 0; j < ncomp; j++) weights[j] = 1.0 / (double)ncomp;
    std::vector<double> inv((size_t)ncomp * c * c), sd(ncomp), post(rows * ncomp), sumw(ncomp);
    double log_lik = 0.0;
    for (int it = 0; it < iters; it++) {
        const int rc = orc_gmm_prepare(covs, ncomp, c, inv.data(), sd.data());
        if (rc != 0) return rc - 100;
        orc_gmm_posteriors(z, rows, c, ncomp, means, inv.data(), sd.data(), weights, post.data());
        (void)log_lik;
        std::fill(sumw.begin(), sumw.end(), 0.0);
        for (size_t r = 0; r < rows; r++)
            for (int j = 0; j < ncomp; j++) sumw[j] += post[r * ncomp + j];
        for (int j = 0; j < ncomp; j++) {
            if (!(sumw[j] > 0.0)) return -2;
            weights[j] = sumw[j] / (double)rows;
            for (int k = 0; k < c; k++) {
                double s = 0.0;
                for (size_t r = 0; r < rows; r++) s += post[r * ncomp + j] * z[r * c + k];
                means[j * c + k] = s / sumw[j];
            }
        }
        for (int j = 0; j < ncomp; j++) {
            double* cv = covs + (size_t)j * c * c;
            std::fill(cv, cv + c * c, 0.0);
            for (size_t r = 0; r < rows; r++) {
                const double w = post[r * ncomp + j];
                for (int a = 0; a < c; a++) {
                    const double da = (z[r * c + a] - means[j * c + a]) * w;
                    for (int b = 0; b < c; b++) cv[a * c + b] += da * (z[r * c + b] - means[j * c + b]);
                }
            }
            for (int k = 0; k < c; k++) cv[k * c + k] += reg;
            for (int k = 0; k < c * c; k++) cv[k] /= sumw[j];
        }
    }
    for (int j = 0; j < ncomp; j++) {  // final check that predict will work
        double det;
        std::vector<double> tmp((size_t)c * c);
        if (!lu_inv_det(covs + (size_t)j * c * c, c, tmp.data(), &det) || !(det > 0.0)) return -3;
    }
    return 0;
}

// ---- row 12: Voting Experts (A7) [RECALL — Cohen & Adams], call sites src/lib.rs:135-136 ----------------------------
// text: n symbols (any byte values). votes: n+1 counters.
void orc_cast_votes(const uint8_t* text, size_t n, int depth, uint32_t* votes) {
    for (size_t i = 0; i <= n; i++) votes[i] = 0;
    if (depth < 1 || n < (size_t)depth) return;
    const int maxlen = depth + 1;
    // n-gram tables per length: count and next-symbol histogram (for boundary entropy)
    struct Node {
        uint32_t count = 0;
        std::map<uint8_t, uint32_t> next;
    };
    std::vector<std::unordered_map<std::string, Node>> tab(maxlen + 1);
    for (int len = 1; len <= maxlen; len++)
        for (size_t s = 0; s + len <= n; s++) {
            Node& nd = tab[len][std::string((const char*)text + s, len)];
            nd.count++;
            if (s + len < n) nd.next[text[s + len]]++;
        }
    // per-length standardisation of frequency and boundary entropy (population std; 0 -> z = 0)
    struct Z {
        double zf, zh;
    };
    std::vector<std::unordered_map<std::string, Z>> zt(maxlen + 1);
    for (int len = 1; len <= maxlen; len++) {
        // deterministic iteration order: sort keys
        std::vector<std::string> keys;
        keys.reserve(tab[len].size());
        for (auto& kv : tab[len]) keys.push_back(kv.first);
        std::sort(keys.begin(), keys.end());
        const double m = (double)keys.size();
        if (keys.empty()) continue;
        std::vector<double> f(keys.size()), h(keys.size());
        for (size_t i = 0; i < keys.size(); i++) {
            const Node& nd = tab[len][keys[i]];
            f[i] = (double)nd.count;
            double tot = 0.0;
            for (auto& kv : nd.next) tot += (double)kv.second;
            double ent = 0.0;
            for (auto& kv : nd.next) {
                const double p = (double)kv.second / tot;
                ent -= p * std::log(p);
            }
            h[i] = ent;
        }
        // frequency statistics from exact integer sums (order-independent, so a parallel implementation reproduces them
        // bit for bit): mean = S1/m, var = S2/m - mean^2. Entropy statistics are plain sequential f64 sums.
        uint64_t s1 = 0, s2 = 0;
        for (size_t i = 0; i < keys.size(); i++) {
            const uint64_t c = (uint64_t)tab[len][keys[i]].count;
            s1 += c;
            s2 += c * c;
        }
        const double mf = (double)s1 / m;
        const double vf = (double)s2 / m - mf * mf;
        double mh = 0;
        for (size_t i = 0; i < keys.size(); i++) mh += h[i];
        mh /= m;
        double vh = 0;
        for (size_t i = 0; i < keys.size(); i++) vh += (h[i] - mh) * (h[i] - mh);
        const double sf = vf > 0 ? std::sqrt(vf) : 0.0, sh = std::sqrt(vh / m);
        for (size_t i = 0; i < keys.size(); i++)
            zt[len][keys[i]] = Z{sf > 0 ? (f[i] - mf) / sf : 0.0, sh > 0 ? (h[i] - mh) / sh : 0.0};
    }
    auto zf = [&](size_t s, int len) { return zt[len].at(std::string((const char*)text + s, len)).zf; };
    auto zh = [&](size_t s, int len) { return zt[len].at(std::string((const char*)text + s, len)).zh; };
    for (size_t s = 0; s + depth <= n; s++) {
        // entropy expert: split p in 1..depth maximising z_H(w[..p]); earliest on ties
        int best_p = 1;
        double best = zh(s, 1);
        for (int p = 2; p <= depth; p++) {
            const double v = zh(s, p);
            if (v > best) best = v, best_p = p;
        }
        votes[s + best_p] += 1;
        // frequency expert: split p in 1..depth-1 maximising z_f(w[..p]) + z_f(w[p..]); earliest on ties
        if (depth >= 2) {
            int bp = 1;
            double bv = zf(s, 1) + zf(s + 1, depth - 1);
            for (int p = 2; p <= depth - 1; p++) {
                const double v = zf(s, p) + zf(s + p, depth - p);
                if (v > bv) bv = v, bp = p;
            }
            votes[s + bp] += 1;
        }
    }
}

// split_string (A7): cut before symbol i (0 < i < n) iff votes[i] > votes[i-1] && votes[i] >= votes[i+1] &&
// votes[i] >= threshold. Writes chunk lengths (in symbols); returns the number of chunks (>= 1 when n > 0).
size_t orc_split(const uint32_t* votes, size_t n, int threshold, uint64_t* out_lens) {
    if (n == 0) return 0;
    size_t nseg = 0, start = 0;
    for (size_t i = 1; i < n; i++) {
        if (votes[i] > votes[i - 1] && votes[i] >= votes[i + 1] && votes[i] >= (uint32_t)threshold) {
            out_lens[nseg++] = i - start;
            start = i;
        }
    }
    out_lens[nseg++] = n - start;
    return nseg;
}

// ---- row 17: clone_from_dictionary sample assembly, src/sound.rs:451-472 + to_sound :475-483 ------------------------
// for target segment t (length tgt_len[t] samples) and matched dictionary sound m = match_idx[t]
// (samples dict_samples[dict_off[m]..dict_off[m+1]]): copy min(len) samples, zero-pad to tgt_len[t]; concatenate.
void orc_resynth(const double* dict_samples, const uint64_t* dict_off, const uint32_t* match_idx, const uint64_t* tgt_len,
                 size_t nseg, double* out) {
    size_t pos = 0;
    for (size_t t = 0; t < nseg; t++) {
        const uint64_t b = dict_off[match_idx[t]], e = dict_off[match_idx[t] + 1];
        const uint64_t have = e - b, want = tgt_len[t];
        const uint64_t ncopy = have < want ? have : want;
        for (uint64_t i = 0; i < ncopy; i++) out[pos + i] = dict_samples[b + i];
        for (uint64_t i = ncopy; i < want; i++) out[pos + i] = 0.0;
        pos += want;
    }
}

int orc_num_threads(void) { return g_threads; }
int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

}  // extern "C"
