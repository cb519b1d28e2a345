// sound.cu — the MFCC subsystem and the sample-domain helpers of the matcher (row 17).
//
//   k_decode_pcm   Sound::from_path's sample conversion                      src/sound.rs:118-120
//   k_mfcc         analyze_mfccs: frame 1024/256, Hann, FFT, 12/13 mel bands, log10, DCT-II x2   src/sound.rs:215-242
//                  (arithmetic of sample::window + vox_box::spectrum::MFCC, [RECALL] A1-A4 in oracle/ASSUMPTIONS.h)
//   k_max_power    analyze_max_power: rectangular 128/64 frames, max RMS         src/sound.rs:244-256 (bit-exact fold)
//   k_colsum_*     analyze_mean_mfccs                                            src/sound.rs:271-286
//   k_resynth      clone_from_dictionary's pad / truncate + to_sound's concat   src/sound.rs:451-483
//
// k_mfcc: one warp per frame, all arithmetic in f64. The real 1024-point transform is done as a 512-point complex
// Stockham FFT (three radix-8 passes, 2 butterflies per lane per pass, the first pass reading the windowed samples
// straight from global memory with coalesced 16-byte loads, the others exchanging through the warp's 8 KB of shared
// memory), followed by the even/odd split for the bins the mel bank needs (2..229 for C = 12 at 44.1 kHz).
#include <cmath>
#include <map>

#include "sound.cuh"

namespace ss {

struct SoundTables {
    double sample_rate = 0;
    int c = 0;
    int bins[SS_MAX_NCOEFFS + 2] = {0};
    DevBuf<double> d_win;     // 1024 Hann weights (A2)
    DevBuf<double2> d_tw;     // exp(-2 pi i j / 1024), j = 0..1023
    DevBuf<double> d_dct;     // c x c: cos(pi k (2n+1) / (2c))
    DevBuf<int> d_bins;       // c + 2
};

struct SoundState {
    std::vector<SoundTables*> tables;
    DevBuf<double> d_samples, d_mfcc, d_partial, d_small;
    DevBuf<int32_t> d_pcm;
    DevBuf<unsigned long long> d_maxbits;
    DevBuf<uint64_t> d_off_a, d_off_b;
    DevBuf<uint32_t> d_idx;
    ~SoundState() {
        for (auto* t : tables) delete t;
    }
};

static void sound_state_free(void* p) { delete static_cast<SoundState*>(p); }

SoundState* sound_state(ss_ctx* ctx) {
    if (!ctx->sound_state) {
        ctx->sound_state = new SoundState();
        ctx->sound_state_free = sound_state_free;
    }
    return static_cast<SoundState*>(ctx->sound_state);
}

static inline double hz_to_mel(double hz) { return 1125.0 * std::log1p(hz / 700.0); }   // [RECALL] vox_box::spectrum
static inline double mel_to_hz(double mel) { return 700.0 * (std::exp(mel / 1125.0) - 1.0); }

static int get_tables(ss_ctx* ctx, double sample_rate, int c, SoundTables** out) {
    SoundState* st = sound_state(ctx);
    for (auto* t : st->tables)
        if (t->sample_rate == sample_rate && t->c == c) {
            *out = t;
            return SS_OK;
        }
    const double kPi = 3.14159265358979323846264338327950288;
    SoundTables* t = new SoundTables();
    t->sample_rate = sample_rate;
    t->c = c;
    // mel band edges (A3b-c): c+2 points spaced (hi-lo)/c in mel, bin = floor((N+1) hz / sr)
    const double lo = hz_to_mel(SS_F_LO), range = hz_to_mel(SS_F_HI) - lo;
    for (int i = 0; i < c + 2; i++) {
        const double point = ((double)i / (double)c) * range + lo;
        t->bins[i] = (int)std::floor((double)(SS_BIN + 1) * mel_to_hz(point) / sample_rate);
    }
    bool ok = t->bins[0] >= 0 && t->bins[c + 1] <= SS_BIN / 2;
    for (int i = 0; i < c + 1; i++) ok = ok && t->bins[i + 1] > t->bins[i];
    if (!ok) {
        delete t;
        return set_error(ctx, SS_ERR_INVALID, "sample rate %.1f Hz: mel band edges leave [0, %d] or collapse (needs sr >= ~17.9 kHz)",
                         sample_rate, SS_BIN / 2);
    }
    std::vector<double> win(SS_BIN), dct((size_t)c * c);
    std::vector<double2> tw(SS_BIN);
    for (int i = 0; i < SS_BIN; i++) {
        win[i] = 0.5 * (1.0 - std::cos(2.0 * kPi * (double)i / (double)(SS_BIN - 1)));  // A2: symmetric Hann
        const double ang = -2.0 * kPi * (double)i / (double)SS_BIN;
        tw[i] = make_double2(std::cos(ang), std::sin(ang));
    }
    for (int k = 0; k < c; k++)
        for (int n = 0; n < c; n++) dct[(size_t)k * c + n] = std::cos(kPi * (double)k * (2.0 * (double)n + 1.0) / (2.0 * (double)c));
    int rc = upload(ctx, t->d_win, win.data(), win.size());
    if (rc == SS_OK) rc = upload(ctx, t->d_tw, tw.data(), tw.size());
    if (rc == SS_OK) rc = upload(ctx, t->d_dct, dct.data(), dct.size());
    if (rc == SS_OK) rc = upload(ctx, t->d_bins, t->bins, (size_t)c + 2);
    if (rc == SS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = set_error(ctx, SS_ERR_CUDA, "table upload failed");
    if (rc != SS_OK) {
        delete t;
        return rc;
    }
    st->tables.push_back(t);
    *out = t;
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void k_decode_pcm(const int32_t* __restrict__ pcm, size_t n, double denom, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ddiv_rn((double)pcm[i], denom);
}
__global__ void k_decode_pcm16(const int16_t* __restrict__ pcm, size_t n, double denom, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ddiv_rn((double)pcm[i], denom);
}

// ---------------------------------------------------------------------------------------------------------------
struct cplx {
    double x, y;
};
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cplx mul_negi(cplx a) { return {a.y, -a.x}; }  // a * (-i)

// natural-order 8-point DFT (decimation in frequency), in place
__device__ __forceinline__ void dft8(cplx* u) {
    const double s = 0.70710678118654752440084436210485;
    cplx a0 = cadd(u[0], u[4]), a1 = cadd(u[1], u[5]), a2 = cadd(u[2], u[6]), a3 = cadd(u[3], u[7]);
    cplx b0 = csub(u[0], u[4]), b1 = csub(u[1], u[5]), b2 = csub(u[2], u[6]), b3 = csub(u[3], u[7]);
    b1 = cmul(b1, cplx{s, -s});
    b2 = mul_negi(b2);
    b3 = cmul(b3, cplx{-s, -s});
    const cplx c0 = cadd(a0, a2), c1 = cadd(a1, a3), c2 = csub(a0, a2), c3 = mul_negi(csub(a1, a3));
    const cplx d0 = cadd(b0, b2), d1 = cadd(b1, b3), d2 = csub(b0, b2), d3 = mul_negi(csub(b1, b3));
    u[0] = cadd(c0, c1);
    u[1] = cadd(d0, d1);
    u[2] = cadd(c2, c3);
    u[3] = cadd(d2, d3);
    u[4] = csub(c0, c1);
    u[5] = csub(d0, d1);
    u[6] = csub(c2, c3);
    u[7] = csub(d2, d3);
}

constexpr int kMfccWarps = 4;
constexpr int kHalf = SS_BIN / 2;  // 512

__global__ void __launch_bounds__(kMfccWarps * 32)
k_mfcc(const double* __restrict__ samples, size_t frames, const double* __restrict__ win, const double2* __restrict__ tw,
       const double* __restrict__ dctm, const int* __restrict__ bins, int c, double energy_floor, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char mfcc_smem[];
    double2* s_buf = reinterpret_cast<double2*>(mfcc_smem);                                   // [warps][512]  32 KB
    double* s_pw = reinterpret_cast<double*>(mfcc_smem + sizeof(double2) * kMfccWarps * kHalf);  // [warps][512]  16 KB
    double* s_le_all = s_pw + kMfccWarps * kHalf;                                             // [warps][16]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* buf = s_buf + warp * kHalf;
    double* pw = s_pw + warp * kHalf;          // power spectrum of the bins the bank reads
    double* s_le = s_le_all + warp * 16;
    const int kb0 = bins[0], kb1 = bins[c + 1];

    for (size_t f = (size_t)blockIdx.x * kMfccWarps + warp; f < frames; f += (size_t)gridDim.x * kMfccWarps) {
        const double2* s2 = reinterpret_cast<const double2*>(samples + f * SS_HOP);
        const double2* w2 = reinterpret_cast<const double2*>(win);
        // ---- pass 0 (Ns = 1): z[n] = (x[2n] w[2n], x[2n+1] w[2n+1]) straight from global memory --------------------
        cplx u[2][8];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int i = lane + 32 * h;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const double2 sv = s2[i + 64 * t];
                const double2 wv = __ldg(&w2[i + 64 * t]);
                u[h][t] = cplx{sv.x * wv.x, sv.y * wv.y};
            }
            dft8(u[h]);
#pragma unroll
            for (int t = 0; t < 8; t++) buf[8 * i + t] = make_double2(u[h][t].x, u[h][t].y);
        }
        __syncwarp();
        // ---- passes 1, 2 (Ns = 8, 64) -------------------------------------------------------------------------------
#pragma unroll
        for (int pass = 1; pass <= 2; pass++) {
            const int p = pass == 1 ? 8 : 64;
            const int tstep = pass == 1 ? 16 : 2;  // twiddle exp(-2 pi i t k / (8p)) = tw[t k 128 / p]
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int i = lane + 32 * h;
                const int k = i & (p - 1);
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const double2 v = buf[i + 64 * t];
                    const double2 w = __ldg(&tw[(t * k * tstep) & (SS_BIN - 1)]);
                    u[h][t] = cmul(cplx{v.x, v.y}, cplx{w.x, w.y});
                }
                dft8(u[h]);
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int i = lane + 32 * h;
                const int k = i & (p - 1);
                const int jo = ((i - k) << 3) + k;
#pragma unroll
                for (int t = 0; t < 8; t++) buf[jo + t * p] = make_double2(u[h][t].x, u[h][t].y);
            }
            __syncwarp();
        }
        // ---- even/odd split: X[k] = E[k] + W_1024^k O[k] for the bins of the mel bank; power |X|^2 (A3) -------------
        for (int k = kb0 + lane; k < kb1; k += 32) {
            const double2 zk = buf[k & (kHalf - 1)];
            const double2 zm = buf[(kHalf - k) & (kHalf - 1)];
            const cplx e = {0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y)};
            const cplx o = {0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x)};  // (Zk - conj(Zm)) / (2i)
            const double2 w = __ldg(&tw[k]);
            const cplx x = cadd(e, cmul(cplx{w.x, w.y}, o));
            pw[k] = x.x * x.x + x.y * x.y;
        }
        __syncwarp();
        // ---- triangular bands (un-normalised; rise starts at 0, fall starts at 1), log10 with the A4 floor ------------
        if (lane < c) {
            const int b0 = bins[lane], b1 = bins[lane + 1], b2 = bins[lane + 2];
            const double up = (double)(b1 - b0), down = (double)(b2 - b1);
            double up_sum = 0.0, down_sum = 0.0;
            for (int k = b0, i = 0; k < b1; k++, i++) up_sum = up_sum + pw[k] * ((double)i / up);
            for (int k = b1, i = 0; k < b2; k++, i++) down_sum = down_sum + pw[k] * (1.0 - (double)i / down);
            const double e = up_sum + down_sum;
            s_le[lane] = log10(e > energy_floor ? e : energy_floor);
        }
        __syncwarp();
        // ---- DCT-II x 2 ---------------------------------------------------------------------------------------------
        if (lane < c) {
            double acc = 0.0;
            for (int n = 0; n < c; n++) acc = acc + s_le[n] * __ldg(&dctm[lane * c + n]);
            out[f * c + lane] = 2.0 * acc;
        }
        __syncwarp();
    }
}

// analyze_max_power: one thread per 128/64 frame, the reference's sequential fold with separate multiply and add
__global__ void k_max_power(const double* __restrict__ samples, size_t frames, unsigned long long* __restrict__ out_bits) {
    const size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double rms = 0.0;
    if (f < frames) {
        const double* s = samples + f * 64;
        double acc = 0.0;
        for (int i = 0; i < 128; i++) acc = __dadd_rn(acc, __dmul_rn(s[i], s[i]));
        rms = __dsqrt_rn(__ddiv_rn(acc, 128.0));
        if (!(rms == rms)) rms = 0.0;  // f64::max ignores NaN (src/sound.rs:255)
    }
    for (int o = 16; o; o >>= 1) rms = fmax(rms, __shfl_xor_sync(0xffffffffu, rms, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(rms));  // rms >= 0
}

// column sums in two deterministic stages
__global__ void k_colsum_partial(const double* __restrict__ m, size_t rows, int c, double* __restrict__ partial) {
    __shared__ double s[256];
    const int col = threadIdx.x % 16, rl = threadIdx.x / 16;  // 16 row-lanes x 16 columns
    double acc = 0.0;
    if (col < c)
        for (size_t r = (size_t)blockIdx.x * 16 + rl; r < rows; r += (size_t)gridDim.x * 16) acc += m[r * c + col];
    s[threadIdx.x] = acc;
    __syncthreads();
    if (rl == 0) {
        for (int j = 1; j < 16; j++) acc += s[j * 16 + col];
        partial[(size_t)blockIdx.x * 16 + col] = acc;
    }
}
__global__ void k_colsum_final(const double* __restrict__ partial, int nblocks, size_t rows, int c, double* __restrict__ out) {
    const int col = threadIdx.x;
    if (col >= c) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; b++) acc += partial[(size_t)b * 16 + col];
    out[col] = acc / (double)rows;  // 0/0 = NaN for an empty sound, as the reference
}

// clone_from_dictionary + to_sound: output sample -> (target segment by binary search) -> dictionary sample or 0
__global__ void k_resynth(const double* __restrict__ dict_samples, const uint64_t* __restrict__ dict_off,
                          const uint32_t* __restrict__ match_idx, const uint64_t* __restrict__ out_off, size_t nseg, size_t total,
                          double* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t lo = 0, hi = nseg;  // last segment whose out_off <= i
        while (hi - lo > 1) {
            const size_t mid = (lo + hi) >> 1;
            if (out_off[mid] <= i) lo = mid;
            else hi = mid;
        }
        const uint64_t within = i - out_off[lo];
        const uint32_t m = match_idx[lo];
        const uint64_t b = dict_off[m], e = dict_off[m + 1];
        out[i] = within < e - b ? dict_samples[b + within] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
static int mfcc_launch(ss_ctx* ctx, const double* d_samples, size_t n, double sample_rate, int c, double* d_out, size_t frames) {
    if (!frames) return SS_OK;
    SoundTables* t = nullptr;
    SS_TRY(get_tables(ctx, sample_rate, c, &t));
    const int grid = (int)std::min<size_t>((frames + kMfccWarps - 1) / kMfccWarps, (size_t)ctx->sm_count * 16);
    const int smem = (int)(sizeof(double2) * kMfccWarps * kHalf + sizeof(double) * kMfccWarps * kHalf + sizeof(double) * kMfccWarps * 16);
    SS_CUDA(ctx, cudaFuncSetAttribute(k_mfcc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_mfcc<<<grid, kMfccWarps * 32, smem, ctx->stream>>>(d_samples, frames, t->d_win.p, t->d_tw.p, t->d_dct.p, t->d_bins.p, c, 1e-10, d_out);
    SS_LAUNCHED(ctx);
    (void)n;
    return SS_OK;
}

static int check_c(ss_ctx* ctx, int c) {
    if (c < 1 || c > SS_MAX_NCOEFFS) return set_error(ctx, SS_ERR_INVALID, "ncoeffs must be in 1..%d (got %d)", SS_MAX_NCOEFFS, c);
    return SS_OK;
}

}  // namespace ss

using namespace ss;

extern "C" {

int ss_decode_pcm(ss_ctx* ctx, const int32_t* pcm, size_t n, int bits, double* out_samples) {
    if (!ctx) return set_error(nullptr, SS_ERR_INVALID, "ctx is NULL");
    if (bits < 1 || bits > 32) return set_error(ctx, SS_ERR_INVALID, "bits_per_sample must be in 1..32 (got %d)", bits);
    if (n && (!pcm || !out_samples)) return set_error(ctx, SS_ERR_INVALID, "NULL buffer");
    if (!n) return SS_OK;
    SS_CUDA(ctx, cudaSetDevice(ctx->device));
    SoundState* st = sound_state(ctx);
    SS_TRY(upload(ctx, st->d_pcm, pcm, n));
    SS_CUDA(ctx, st->d_samples.reserve(n));
    const double denom = (double)(INT32_MAX >> (32 - bits));  // i32::max_value().wrapping_shr(32 - bits)
    k_decode_pcm<<<ceil_div((long long)n, 256), 256, 0, ctx->stream>>>(st->d_pcm.p, n, denom, st->d_samples.p);
    SS_LAUNCHED(ctx);
    SS_CUDA(ctx, cudaMemcpyAsync(out_samples, st->d_samples.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int ss_mfcc_dev(ss_ctx* ctx, const double* d_samples, size_t n, double sample_rate, int ncoeffs, double* d_out_mfcc) {
    if (!ctx) return set_error(nullptr, SS_ERR_INVALID, "ctx is NULL");
    SS_TRY(check_c(ctx, ncoeffs));
    size_t frames = 0;
    ss_frame_count(n, &frames);
    if (frames && (!d_samples || !d_out_mfcc)) return set_error(ctx, SS_ERR_INVALID, "NULL buffer");
    if (((uintptr_t)d_samples & 15) != 0) return set_error(ctx, SS_ERR_INVALID, "d_samples must be 16-byte aligned");
    SS_CUDA(ctx, cudaSetDevice(ctx->device));
    return mfcc_launch(ctx, d_samples, n, sample_rate, ncoeffs, d_out_mfcc, frames);
}

// the three analyses on samples already in st->d_samples; stream-synchronised on return
static int analyze_resident(ss_ctx* ctx, SoundState* st, size_t n, double sample_rate, int ncoeffs, double* out_mfcc, size_t* out_frames,
                            double* out_max_power, double* out_mean_mfccs) {
    size_t frames = 0;
    ss_frame_count(n, &frames);
    if (out_frames) *out_frames = frames;
    const bool want_mfcc = out_mfcc || out_mean_mfccs;
    if (want_mfcc && frames) {
        SS_CUDA(ctx, st->d_mfcc.reserve(frames * (size_t)ncoeffs));
        SS_TRY(mfcc_launch(ctx, st->d_samples.p, n, sample_rate, ncoeffs, st->d_mfcc.p, frames));
        if (out_mfcc)
            SS_CUDA(ctx, cudaMemcpyAsync(out_mfcc, st->d_mfcc.p, frames * (size_t)ncoeffs * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (out_mean_mfccs) {
        SS_CUDA(ctx, st->d_small.reserve(16));
        const int nb = (int)std::min<size_t>(std::max<size_t>(frames / 64, 1), 1024);
        SS_CUDA(ctx, st->d_partial.reserve((size_t)nb * 16));
        k_colsum_partial<<<nb, 256, 0, ctx->stream>>>(st->d_mfcc.p, frames, ncoeffs, st->d_partial.p);
        SS_LAUNCHED(ctx);
        k_colsum_final<<<1, 32, 0, ctx->stream>>>(st->d_partial.p, nb, frames, ncoeffs, st->d_small.p);
        SS_LAUNCHED(ctx);
        SS_CUDA(ctx, cudaMemcpyAsync(out_mean_mfccs, st->d_small.p, (size_t)ncoeffs * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    unsigned long long bits = 0;
    if (out_max_power) {
        SS_CUDA(ctx, st->d_maxbits.reserve(1));
        SS_CUDA(ctx, cudaMemsetAsync(st->d_maxbits.p, 0, sizeof(unsigned long long), ctx->stream));
        const size_t pframes = n >= 128 ? (n - 128) / 64 + 1 : 0;
        if (pframes) {
            k_max_power<<<ceil_div((long long)pframes, 128), 128, 0, ctx->stream>>>(st->d_samples.p, pframes, st->d_maxbits.p);
            SS_LAUNCHED(ctx);
        }
        SS_CUDA(ctx, cudaMemcpyAsync(&bits, st->d_maxbits.p, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    }
    SS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_max_power) memcpy(out_max_power, &bits, sizeof(double));
    return SS_OK;
}

int ss_sound_analyze(ss_ctx* ctx, const double* samples, size_t n, double sample_rate, int ncoeffs, double* out_mfcc,
                     size_t* out_frames, double* out_max_power, double* out_mean_mfccs) {
    if (!ctx) return set_error(nullptr, SS_ERR_INVALID, "ctx is NULL");
    SS_TRY(check_c(ctx, ncoeffs));
    if (n && !samples) return set_error(ctx, SS_ERR_INVALID, "samples is NULL");
    SS_CUDA(ctx, cudaSetDevice(ctx->device));
    SoundState* st = sound_state(ctx);
    SS_TRY(upload(ctx, st->d_samples, samples, n));
    return analyze_resident(ctx, st, n, sample_rate, ncoeffs, out_mfcc, out_frames, out_max_power, out_mean_mfccs);
}

int ss_sound_analyze_pcm(ss_ctx* ctx, const void* pcm, size_t n, int bits_per_sample, double sample_rate, int ncoeffs, double* out_samples,
                         double* out_mfcc, size_t* out_frames, double* out_max_power, double* out_mean_mfccs) {
    if (!ctx) return set_error(nullptr, SS_ERR_INVALID, "ctx is NULL");
    SS_TRY(check_c(ctx, ncoeffs));
    if (bits_per_sample != 16 && bits_per_sample != 24 && bits_per_sample != 32)
        return set_error(ctx, SS_ERR_INVALID, "bits_per_sample must be 16 (int16 buffer), 24 or 32 (int32 buffer); got %d", bits_per_sample);
    if (n && !pcm) return set_error(ctx, SS_ERR_INVALID, "pcm is NULL");
    SS_CUDA(ctx, cudaSetDevice(ctx->device));
    SoundState* st = sound_state(ctx);
    SS_CUDA(ctx, st->d_samples.reserve(n));
    const double denom = (double)(INT32_MAX >> (32 - bits_per_sample));  // src/sound.rs:118-120
    if (n) {
        if (bits_per_sample == 16) {
            SS_CUDA(ctx, st->d_pcm.reserve((n + 1) / 2));  // int16 samples packed in the int32 staging buffer
            SS_CUDA(ctx, cudaMemcpyAsync(st->d_pcm.p, pcm, n * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
            k_decode_pcm16<<<ceil_div((long long)n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const int16_t*>(st->d_pcm.p), n, denom,
                                                                               st->d_samples.p);
        } else {
            SS_TRY(upload(ctx, st->d_pcm, static_cast<const int32_t*>(pcm), n));
            k_decode_pcm<<<ceil_div((long long)n, 256), 256, 0, ctx->stream>>>(st->d_pcm.p, n, denom, st->d_samples.p);
        }
        SS_LAUNCHED(ctx);
        if (out_samples) SS_CUDA(ctx, cudaMemcpyAsync(out_samples, st->d_samples.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    return analyze_resident(ctx, st, n, sample_rate, ncoeffs, out_mfcc, out_frames, out_max_power, out_mean_mfccs);
}

int ss_mfcc(ss_ctx* ctx, const double* samples, size_t n, double sample_rate, int ncoeffs, double* out_mfcc, size_t* out_frames) {
    return ss_sound_analyze(ctx, samples, n, sample_rate, ncoeffs, out_mfcc, out_frames, nullptr, nullptr);
}

int ss_max_power(ss_ctx* ctx, const double* samples, size_t n, double* out) {
    if (!out) return set_error(ctx, SS_ERR_INVALID, "out is NULL");
    return ss_sound_analyze(ctx, samples, n, 44100.0, SS_NCOEFFS, nullptr, nullptr, out, nullptr);
}

int ss_resynth(ss_ctx* ctx, const double* dict_samples, const uint64_t* dict_sample_offsets, size_t ndict, const uint32_t* match_idx,
               const uint64_t* target_lens, size_t nseg, double* out_samples) {
    if (!ctx) return set_error(nullptr, SS_ERR_INVALID, "ctx is NULL");
    if (!nseg) return SS_OK;
    if (!dict_sample_offsets || !match_idx || !target_lens || !out_samples) return set_error(ctx, SS_ERR_INVALID, "NULL buffer");
    if (!ndict) return set_error(ctx, SS_ERR_EMPTY_DICT, "resynthesis from an empty dictionary");
    for (size_t i = 0; i < ndict; i++)
        if (dict_sample_offsets[i + 1] < dict_sample_offsets[i]) return set_error(ctx, SS_ERR_INVALID, "dictionary offsets not monotone at %zu", i);
    std::vector<uint64_t> out_off(nseg + 1, 0);
    for (size_t i = 0; i < nseg; i++) {
        if (match_idx[i] >= ndict) return set_error(ctx, SS_ERR_INVALID, "match_idx[%zu] = %u out of range", i, match_idx[i]);
        out_off[i + 1] = out_off[i] + target_lens[i];
    }
    const size_t total = out_off[nseg], dtotal = dict_sample_offsets[ndict];
    if (!total) return SS_OK;
    if (dtotal && !dict_samples) return set_error(ctx, SS_ERR_INVALID, "dict_samples is NULL");
    SS_CUDA(ctx, cudaSetDevice(ctx->device));
    SoundState* st = sound_state(ctx);
    SS_TRY(upload(ctx, st->d_samples, dict_samples, dtotal));
    SS_TRY(upload(ctx, st->d_off_a, dict_sample_offsets, ndict + 1));
    SS_TRY(upload(ctx, st->d_off_b, out_off.data(), nseg + 1));
    SS_TRY(upload(ctx, st->d_idx, match_idx, nseg));
    SS_CUDA(ctx, st->d_mfcc.reserve(total));
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
    k_resynth<<<grid, 256, 0, ctx->stream>>>(st->d_samples.p, st->d_off_a.p, st->d_idx.p, st->d_off_b.p, nseg, total, st->d_mfcc.p);
    SS_LAUNCHED(ctx);
    SS_CUDA(ctx, cudaMemcpyAsync(out_samples, st->d_mfcc.p, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

}  // extern "C"
