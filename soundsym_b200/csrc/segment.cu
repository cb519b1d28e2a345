// segment.cu — the segmentation subsystem and SoundSequence::new's distance list (row 18).
// Compiled with --fmad=false: the GMM quadratic forms and the z-scores follow the CPU path operation for operation.
//
//   ss_symbols    discretize_with_model (src/lib.rs:56-60): Standardizer re-fit on this input (A5), GMM posteriors (A6),
//                 then max_index per row (src/lib.rs:123-131, src/sound.rs:486-495) -> 'A' + idx
//   ss_vote_split voting_experts::cast_votes / split_string (src/lib.rs:135-137), spec A7:
//                 packed n-gram keys -> radix sort -> run-length encode (counts) -> per-length z-scores of frequency
//                 (exact integer sums) and boundary entropy -> one thread per window casts the two experts' votes ->
//                 local-maximum / threshold flags -> stream compaction -> segment lengths x HOP
//   ss_sequence_distances  cosine_sim_angular over consecutive mean-MFCC rows (src/sound.rs:62-69, 392-396)
#include <cub/cub.cuh>

#include <algorithm>
#include <random>

#include "sound.cuh"

namespace ss {

constexpr int kMaxComp = 32;
constexpr int kMaxDepth = 7;  // n-gram keys of depth+1 bytes are packed into one u64

struct SegState {
    DevBuf<double> d_mfcc, d_z, d_post, d_model, d_inv, d_sqrt_det, d_partial, d_stats;
    DevBuf<uint8_t> d_sym;
    DevBuf<int> d_status;
    // voting experts
    DevBuf<unsigned long long> d_keys, d_keys_sorted;
    DevBuf<unsigned long long> d_uniq[kMaxDepth + 2];
    DevBuf<uint32_t> d_cnt[kMaxDepth + 2];
    DevBuf<double> d_zf[kMaxDepth + 2], d_zh[kMaxDepth + 2], d_h;
    DevBuf<uint32_t> d_rank[kMaxDepth + 2];
    DevBuf<uint32_t> d_votes, d_nruns, d_bounds, d_nbounds;
    DevBuf<uint8_t> d_flags, d_cub_tmp;
    DevBuf<unsigned long long> d_isums;
    DevBuf<uint64_t> d_lens;
};
static void seg_state_free(void* p) { delete static_cast<SegState*>(p); }
static SegState* seg_state(ss_ctx* ctx) {
    if (!ctx->seg_state) {
        ctx->seg_state = new SegState();
        ctx->seg_state_free = seg_state_free;
    }
    return static_cast<SegState*>(ctx->seg_state);
}

// ---------------------------------------------------------------------------------------------------------------
// Standardizer (A5): per-column mean and (n-1) variance, deterministic two-stage tree
// ---------------------------------------------------------------------------------------------------------------
// stage 1: block b accumulates rows b, b+grid, ... ; which == 0: sum x ; which == 1: sum (mean - x)^2
__global__ void k_col_partial(const double* __restrict__ x, size_t rows, int c, const double* __restrict__ mean, int which,
                              double* __restrict__ partial) {
    __shared__ double s[256];
    const int col = threadIdx.x % 16, rl = threadIdx.x / 16;
    double acc = 0.0;
    if (col < c) {
        const double mu = which ? mean[col] : 0.0;
        for (size_t r = (size_t)blockIdx.x * 16 + rl; r < rows; r += (size_t)gridDim.x * 16) {
            const double v = x[r * c + col];
            if (which) {
                const double d = mu - v;
                acc += d * d;
            } else {
                acc += v;
            }
        }
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    if (rl == 0) {
        for (int j = 1; j < 16; j++) acc += s[j * 16 + col];
        partial[(size_t)blockIdx.x * 16 + col] = acc;
    }
}
// stage 2: out[col] = (sum of partials) / denom ; sqrt_out (optional) = sqrt of that
__global__ void k_col_final(const double* __restrict__ partial, int nblocks, int c, double denom, double* __restrict__ out,
                            double* __restrict__ sqrt_out) {
    const int col = threadIdx.x;
    if (col >= c) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; b++) acc += partial[(size_t)b * 16 + col];
    const double v = acc / denom;
    out[col] = v;
    if (sqrt_out) sqrt_out[col] = sqrt(v);
}
__global__ void k_standardize(const double* __restrict__ x, size_t rows, int c, const double* __restrict__ mean,
                              const double* __restrict__ sd, double* __restrict__ z) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * (size_t)c) return;
    const int col = (int)(i % c);
    z[i] = (x[i] - mean[col]) / sd[col];
}

// ---------------------------------------------------------------------------------------------------------------
// GMM (A6): inverse and determinant of every covariance by LU with partial pivoting (one thread per component, the
// CPU path's operation order), then one thread per frame evaluates the 26 quadratic forms.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_gmm_prepare(const double* __restrict__ covs, int ncomp, int c, double* __restrict__ inv, double* __restrict__ sqrt_det,
                              int* __restrict__ status) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncomp) return;
    double lu[SS_MAX_NCOEFFS * SS_MAX_NCOEFFS];
    int perm[SS_MAX_NCOEFFS];
    const double* a = covs + (size_t)j * c * c;
    for (int i = 0; i < c * c; i++) lu[i] = a[i];
    for (int i = 0; i < c; i++) perm[i] = i;
    double d = 1.0;
    for (int col = 0; col < c; col++) {
        int piv = col;
        double best = fabs(lu[col * c + col]);
        for (int r = col + 1; r < c; r++)
            if (fabs(lu[r * c + col]) > best) best = fabs(lu[r * c + col]), piv = r;
        if (best == 0.0) {
            atomicMin(status, -(j + 1));
            return;
        }
        if (piv != col) {
            for (int k = 0; k < c; k++) {
                const double t = lu[piv * c + k];
                lu[piv * c + k] = lu[col * c + k];
                lu[col * c + k] = t;
            }
            const int tp = perm[piv];
            perm[piv] = perm[col];
            perm[col] = tp;
            d = -d;
        }
        d *= lu[col * c + col];
        for (int r = col + 1; r < c; r++) {
            lu[r * c + col] /= lu[col * c + col];
            const double f = lu[r * c + col];
            for (int k = col + 1; k < c; k++) lu[r * c + k] -= f * lu[col * c + k];
        }
    }
    sqrt_det[j] = sqrt(d);
    double* out = inv + (size_t)j * c * c;
    double y[SS_MAX_NCOEFFS];
    for (int e = 0; e < c; e++) {
        for (int i = 0; i < c; i++) {
            double s = perm[i] == e ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) s -= lu[i * c + k] * y[k];
            y[i] = s;
        }
        for (int i = c - 1; i >= 0; i--) {
            double s = y[i];
            for (int k = i + 1; k < c; k++) s -= lu[i * c + k] * out[k * c + e];
            out[i * c + e] = s / lu[i * c + i];
        }
    }
}

__global__ void __launch_bounds__(128)
k_gmm_symbols(const double* __restrict__ z, size_t rows, int c, int ncomp, const double* __restrict__ means, const double* __restrict__ inv,
              const double* __restrict__ sqrt_det, const double* __restrict__ weights, uint8_t* __restrict__ sym,
              double* __restrict__ post) {
    extern __shared__ double sm[];
    double* s_inv = sm;                               // ncomp * c * c
    double* s_mean = s_inv + (size_t)ncomp * c * c;   // ncomp * c
    double* s_sd = s_mean + (size_t)ncomp * c;        // ncomp
    double* s_w = s_sd + ncomp;                       // ncomp
    for (int i = threadIdx.x; i < ncomp * c * c; i += blockDim.x) s_inv[i] = inv[i];
    for (int i = threadIdx.x; i < ncomp * c; i += blockDim.x) s_mean[i] = means[i];
    for (int i = threadIdx.x; i < ncomp; i += blockDim.x) s_sd[i] = sqrt_det[i], s_w[i] = weights[i];
    __syncthreads();
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double x[SS_MAX_NCOEFFS];
    for (int k = 0; k < c; k++) x[k] = z[r * c + k];
    double pdfs[kMaxComp];
    for (int j = 0; j < ncomp; j++) {
        const double* P = s_inv + (size_t)j * c * c;
        const double* mu = s_mean + (size_t)j * c;
        double quad = 0.0;
        for (int col = 0; col < c; col++) {  // (diff * P) * diff^T, as `&diff * &cov_inv * diff.transpose()`
            double s = 0.0;
            for (int k = 0; k < c; k++) s = s + (x[k] - mu[k]) * P[k * c + col];
            quad = quad + s * (x[col] - mu[col]);
        }
        pdfs[j] = exp(quad * -0.5) / s_sd[j];
    }
    // weighted sum with rulinalg's dot (A9)
    double p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int e = 0;
    for (; e + 8 <= ncomp; e += 8)
        for (int u = 0; u < 8; u++) p[u] = p[u] + s_w[e + u] * pdfs[e + u];
    double wsum = 0.0;
    wsum = wsum + (p[0] + p[4]);
    wsum = wsum + (p[1] + p[5]);
    wsum = wsum + (p[2] + p[6]);
    wsum = wsum + (p[3] + p[7]);
    for (; e < ncomp; e++) wsum = wsum + s_w[e] * pdfs[e];
    // max_index: first index whose value is > the running max starting from (0, 0.0)
    int best = 0;
    double bestv = 0.0;
    for (int j = 0; j < ncomp; j++) {
        const double w = s_w[j] * pdfs[j] / wsum;
        if (post) post[r * ncomp + j] = w;
        if (w > bestv) bestv = w, best = j;
    }
    sym[r] = (uint8_t)('A' + best);
}

// ---------------------------------------------------------------------------------------------------------------
// Voting Experts (A7)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_key(const uint8_t* __restrict__ text, size_t s, int len) {
    unsigned long long k = 0;
    for (int i = 0; i < len; i++) k = (k << 8) | text[s + i];  // big-endian: integer order == lexicographic order
    return k;
}
__global__ void k_ve_keys(const uint8_t* __restrict__ text, size_t n, int len, unsigned long long* __restrict__ keys) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s + len <= n) keys[s] = pack_key(text, s, len);
}
// exact integer sums of counts and squared counts over the unique n-grams of one length
__global__ void k_ve_isums(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ nruns, unsigned long long* __restrict__ sums) {
    const uint32_t m = *nruns;
    unsigned long long s1 = 0, s2 = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const unsigned long long c = cnt[i];
        s1 += c;
        s2 += c * c;
    }
    for (int o = 16; o; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[0], s1);
        atomicAdd(&sums[1], s2);
    }
}
__global__ void k_ve_zfreq(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ nruns, const unsigned long long* __restrict__ sums,
                           double* __restrict__ zf) {
    const uint32_t m = *nruns;
    const double md = (double)m;
    const double mf = (double)sums[0] / md;
    const double vf = (double)sums[1] / md - mf * mf;
    const double sf = vf > 0 ? sqrt(vf) : 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
        zf[i] = sf > 0 ? ((double)cnt[i] - mf) / sf : 0.0;
}
// boundary entropy of every unique len-gram from the sorted unique (len+1)-grams (grouped by prefix, symbols ascending)
__global__ void k_ve_entropy(const unsigned long long* __restrict__ uniq, const uint32_t* __restrict__ nruns,
                             const unsigned long long* __restrict__ uniq_next, const uint32_t* __restrict__ cnt_next,
                             const uint32_t* __restrict__ nruns_next, double* __restrict__ h) {
    const uint32_t m = *nruns, mn = nruns_next ? *nruns_next : 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const unsigned long long lo_key = uniq[i] << 8;
        uint32_t lo = 0, hi = mn;  // first index with uniq_next >= lo_key
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (uniq_next[mid] < lo_key) lo = mid + 1;
            else hi = mid;
        }
        const uint32_t first = lo;
        double tot = 0.0;
        uint32_t e = first;
        for (; e < mn && (uniq_next[e] >> 8) == uniq[i]; e++) tot += (double)cnt_next[e];
        double ent = 0.0;
        for (uint32_t q = first; q < e; q++) {
            const double p = (double)cnt_next[q] / tot;
            ent -= p * log(p);
        }
        h[i] = ent;
    }
}
// mean / population std of h over the unique n-grams: single block, fixed-order tree -> stats[0] = mean, stats[1] = std
__global__ void k_ve_hstats(const double* __restrict__ h, const uint32_t* __restrict__ nruns, double* __restrict__ stats) {
    __shared__ double s[256];
    const uint32_t m = *nruns;
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < m; i += 256) acc += h[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    const double mh = s[0] / (double)m;
    __syncthreads();
    acc = 0.0;
    for (uint32_t i = threadIdx.x; i < m; i += 256) acc += (h[i] - mh) * (h[i] - mh);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[0] = mh;
        stats[1] = sqrt(s[0] / (double)m);
    }
}
__global__ void k_ve_zent(const double* __restrict__ h, const uint32_t* __restrict__ nruns, const double* __restrict__ stats,
                          double* __restrict__ zh) {
    const uint32_t m = *nruns;
    const double mh = stats[0], sh = stats[1];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) zh[i] = sh > 0 ? (h[i] - mh) / sh : 0.0;
}
// rank[s] = index of the len-gram starting at s among the sorted unique len-grams
__global__ void k_ve_rank(const uint8_t* __restrict__ text, size_t n, int len, const unsigned long long* __restrict__ uniq,
                          const uint32_t* __restrict__ nruns, uint32_t* __restrict__ rank) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s + len > n) return;
    const unsigned long long key = pack_key(text, s, len);
    uint32_t lo = 0, hi = *nruns;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (uniq[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    rank[s] = lo;
}
struct VeTables {
    const double* zf[kMaxDepth + 2];
    const double* zh[kMaxDepth + 2];
    const uint32_t* rank[kMaxDepth + 2];
};
// one thread per window position: the entropy expert and the frequency expert each add one vote
__global__ void k_ve_votes(size_t n, int depth, VeTables t, uint32_t* __restrict__ votes) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s + depth > n) return;
    int best_p = 1;
    double best = t.zh[1][t.rank[1][s]];
    for (int p = 2; p <= depth; p++) {
        const double v = t.zh[p][t.rank[p][s]];
        if (v > best) best = v, best_p = p;
    }
    atomicAdd(&votes[s + best_p], 1u);
    if (depth >= 2) {
        int bp = 1;
        double bv = t.zf[1][t.rank[1][s]] + t.zf[depth - 1][t.rank[depth - 1][s + 1]];
        for (int p = 2; p <= depth - 1; p++) {
            const double v = t.zf[p][t.rank[p][s]] + t.zf[depth - p][t.rank[depth - p][s + p]];
            if (v > bv) bv = v, bp = p;
        }
        atomicAdd(&votes[s + bp], 1u);
    }
}
__global__ void k_ve_flags(const uint32_t* __restrict__ votes, size_t n, uint32_t threshold, uint8_t* __restrict__ flags) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i > 0 && votes[i] > votes[i - 1] && votes[i] >= votes[i + 1] && votes[i] >= threshold) ? 1 : 0;
}
// boundaries (sorted cut positions) -> chunk lengths in samples
__global__ void k_ve_lens(const uint32_t* __restrict__ bounds, const uint32_t* __restrict__ nb, size_t n, uint64_t hop,
                          uint64_t* __restrict__ lens) {
    const uint32_t m = *nb;  // chunks = m + 1
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= m; i += gridDim.x * blockDim.x) {
        const uint64_t b = i ? bounds[i - 1] : 0, e = i < m ? bounds[i] : (uint64_t)n;
        lens[i] = (e - b) * hop;
    }
}

// cosine_sim_angular of consecutive rows (src/sound.rs:62-69): norm = sum of squares (no sqrt), dot per A9
__global__ void k_seq_dist(const double* __restrict__ m, size_t nrows, int c, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= nrows) return;
    const double* a = m + i * c;
    const double* b = a + c;
    double na = 0.0, nb = 0.0;
    for (int k = 0; k < c; k++) na = a[k] * a[k] + na;
    for (int k = 0; k < c; k++) nb = b[k] * b[k] + nb;
    double p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int e = 0;
    for (; e + 8 <= c; e += 8)
        for (int u = 0; u < 8; u++) p[u] = p[u] + a[e + u] * b[e + u];
    double s = 0.0;
    s = s + (p[0] + p[4]);
    s = s + (p[1] + p[5]);
    s = s + (p[2] + p[6]);
    s = s + (p[3] + p[7]);
    for (; e < c; e++) s = s + a[e] * b[e];
    double sim = s / (na * nb);
    if (sim > 1.0) sim = 1.0;
    else if (sim < -1.0) sim = 1.0;  // sic, src/sound.rs:65
    out[i] = acos(sim) * 0.318309886183790671537767526745028724;
}

// ---------------------------------------------------------------------------------------------------------------
static int symbols_dev(ss_ctx* ctx, SegState* st, const double* d_mfcc, size_t frames, const ss_gmm* model, uint8_t* d_sym,
                       double* d_post) {
    const int c = model->ncoeffs, nc = model->ncomp;
    const size_t msz = (size_t)nc * c + (size_t)nc * c * c + nc;
    std::vector<double> hm(msz);
    memcpy(hm.data(), model->means, sizeof(double) * nc * c);
    memcpy(hm.data() + (size_t)nc * c, model->covs, sizeof(double) * nc * c * c);
    memcpy(hm.data() + (size_t)nc * c + (size_t)nc * c * c, model->weights, sizeof(double) * nc);
    SS_TRY(upload(ctx, st->d_model, hm.data(), msz));
    const double* d_means = st->d_model.p;
    const double* d_covs = d_means + (size_t)nc * c;
    const double* d_w = d_covs + (size_t)nc * c * c;
    SS_CUDA(ctx, st->d_inv.reserve((size_t)nc * c * c));
    SS_CUDA(ctx, st->d_sqrt_det.reserve(nc));
    SS_CUDA(ctx, st->d_status.reserve(1));
    SS_CUDA(ctx, cudaMemsetAsync(st->d_status.p, 0, sizeof(int), ctx->stream));
    k_gmm_prepare<<<1, 32, 0, ctx->stream>>>(d_covs, nc, c, st->d_inv.p, st->d_sqrt_det.p, st->d_status.p);
    SS_LAUNCHED(ctx);
    // Standardizer
    const int nb = (int)std::min<size_t>(std::max<size_t>(frames / 64, 1), 1024);
    SS_CUDA(ctx, st->d_partial.reserve((size_t)nb * 16));
    SS_CUDA(ctx, st->d_stats.reserve(48));
    double* d_mean = st->d_stats.p;
    double* d_var = d_mean + 16;
    double* d_sd = d_mean + 32;
    k_col_partial<<<nb, 256, 0, ctx->stream>>>(d_mfcc, frames, c, nullptr, 0, st->d_partial.p);
    SS_LAUNCHED(ctx);
    k_col_final<<<1, 32, 0, ctx->stream>>>(st->d_partial.p, nb, c, (double)frames, d_mean, nullptr);
    SS_LAUNCHED(ctx);
    k_col_partial<<<nb, 256, 0, ctx->stream>>>(d_mfcc, frames, c, d_mean, 1, st->d_partial.p);
    SS_LAUNCHED(ctx);
    k_col_final<<<1, 32, 0, ctx->stream>>>(st->d_partial.p, nb, c, (double)(frames - 1), d_var, d_sd);
    SS_LAUNCHED(ctx);
    SS_CUDA(ctx, st->d_z.reserve(frames * (size_t)c));
    k_standardize<<<ceil_div((long long)(frames * c), 256), 256, 0, ctx->stream>>>(d_mfcc, frames, c, d_mean, d_sd, st->d_z.p);
    SS_LAUNCHED(ctx);
    const size_t smem = sizeof(double) * ((size_t)nc * c * c + (size_t)nc * c + 2 * (size_t)nc);
    SS_CUDA(ctx, cudaFuncSetAttribute(k_gmm_symbols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gmm_symbols<<<ceil_div((long long)frames, 128), 128, smem, ctx->stream>>>(st->d_z.p, frames, c, nc, d_means, st->d_inv.p,
                                                                              st->d_sqrt_det.p, d_w, d_sym, d_post);
    SS_LAUNCHED(ctx);
    // the host staging vector is read by the async upload: wait before it goes out of scope
    int status = 0;
    SS_CUDA(ctx, cudaMemcpyAsync(&status, st->d_status.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (status != 0) return set_error(ctx, SS_ERR_INVALID, "covariance of component %d is singular", -status - 1);
    return SS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// GMM training (train_model, src/lib.rs:44-54; EM of rusty-machine 0.5.4 [RECALL], restated in the oracle's orc_gmm_train)
// ---------------------------------------------------------------------------------------------------------------
// weighted moments of one component in two deterministic stages. Block (b, j) walks rows b, b+gridDim.x, ... in tiles of
// 32 staged in shared memory; thread t < c*c owns covariance element (t / c, t % c), threads c*c .. c*c+c-1 own the
// weighted sums of z, thread c*c+c owns the weight sum.  mode 0: w = post[r][j], centre = 0 (first pass: sums for the
// means); mode 1: w = post[r][j], centre = means[j] (second pass: covariance); mode 2: w = 1, centre = centre0 (initial
// data covariance).
__global__ void __launch_bounds__(256)
k_gmm_moments(const double* __restrict__ z, size_t rows, int c, int ncomp, const double* __restrict__ post, const double* __restrict__ centre,
              int mode, double* __restrict__ partial) {
    __shared__ double sz[32][SS_MAX_NCOEFFS];
    __shared__ double sw[32];
    const int j = blockIdx.y, t = threadIdx.x;
    const int nout = c * c + c + 1;
    const int a = t / c, b2 = t % c;
    double mu_a = 0.0, mu_b = 0.0;
    if (mode != 0 && t < c * c) {
        const double* mu = centre + (mode == 1 ? (size_t)j * c : 0);
        mu_a = mu[a];
        mu_b = mu[b2];
    }
    double acc = 0.0;
    for (size_t r0 = (size_t)blockIdx.x * 32; r0 < rows; r0 += (size_t)gridDim.x * 32) {
        __syncthreads();
        for (int i = t; i < 32 * c; i += blockDim.x) {
            const size_t r = r0 + i / c;
            sz[i / c][i % c] = r < rows ? z[r * c + i % c] : 0.0;
        }
        if (t < 32) {
            const size_t r = r0 + t;
            sw[t] = r < rows ? (mode == 2 ? 1.0 : post[r * ncomp + j]) : 0.0;
        }
        __syncthreads();
        if (t < c * c) {
            for (int i = 0; i < 32; i++) acc += ((sz[i][a] - mu_a) * sw[i]) * (sz[i][b2] - mu_b);
        } else if (t < c * c + c) {
            const int k = t - c * c;
            for (int i = 0; i < 32; i++) acc += sw[i] * sz[i][k];
        } else if (t == c * c + c) {
            for (int i = 0; i < 32; i++) acc += sw[i];
        }
    }
    if (t < nout) partial[((size_t)j * gridDim.x + blockIdx.x) * nout + t] = acc;
}
// sums the per-block partials in block order -> out[j][nout]
__global__ void k_gmm_moments_final(const double* __restrict__ partial, int nblocks, int nout, double* __restrict__ out) {
    const int j = blockIdx.x, t = threadIdx.x;
    if (t >= nout) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; b++) acc += partial[((size_t)j * nblocks + b) * nout + t];
    out[(size_t)j * nout + t] = acc;
}
// M-step, part 1: weights and means from the first-pass moments
__global__ void k_gmm_update_means(const double* __restrict__ mom, int c, int ncomp, size_t rows, double* __restrict__ means,
                                   double* __restrict__ weights, int* __restrict__ status) {
    const int j = blockIdx.x, k = threadIdx.x;
    const int nout = c * c + c + 1;
    const double sumw = mom[(size_t)j * nout + c * c + c];
    if (k == 0) {
        weights[j] = sumw / (double)rows;
        if (!(sumw > 0.0)) atomicMin(status, -1000 - j);
    }
    if (k < c) means[(size_t)j * c + k] = mom[(size_t)j * nout + c * c + k] / sumw;
}
// M-step, part 2: cov = (sum_i w (z - mu)(z - mu)^T + reg I) / sum w   (Regularized(reg), [RECALL])
__global__ void k_gmm_update_covs(const double* __restrict__ mom, int c, double reg, double* __restrict__ covs) {
    const int j = blockIdx.x, t = threadIdx.x;
    const int nout = c * c + c + 1;
    if (t >= c * c) return;
    const double sumw = mom[(size_t)j * nout + c * c + c];
    double v = mom[(size_t)j * nout + t];
    if (t / c == t % c) v += reg;
    covs[(size_t)j * c * c + t] = v / sumw;
}
// initial covariance for every component: data covariance / (n - 1) + reg I
__global__ void k_gmm_init_covs(const double* __restrict__ mom, int c, int ncomp, size_t rows, double reg, double* __restrict__ covs) {
    const int j = blockIdx.x, t = threadIdx.x;
    if (t >= c * c) return;
    double v = mom[t] * (1.0 / (double)(rows - 1));
    if (t / c == t % c) v += reg;
    covs[(size_t)j * c * c + t] = v;
}
__global__ void k_gmm_init_means(const double* __restrict__ z, int c, const uint64_t* __restrict__ pick, double* __restrict__ means,
                                 double* __restrict__ weights, int ncomp) {
    const int j = blockIdx.x, k = threadIdx.x;
    if (k < c) means[(size_t)j * c + k] = z[pick[j] * c + k];
    if (k == 0) weights[j] = 1.0 / (double)ncomp;
}
__global__ void k_gmm_check_det(const double* __restrict__ sqrt_det, int ncomp, int* __restrict__ status) {
    const int j = threadIdx.x;
    if (j < ncomp && !(sqrt_det[j] > 0.0)) atomicMin(status, -2000 - j);  // sqrt of a non-positive determinant is NaN / 0
}

static int check_model(ss_ctx* ctx, const ss_gmm* model) {
    if (!model) return set_error(ctx, SS_ERR_NOT_TRAINED, "Must first train model");  // src/lib.rs:141
    if (model->ncomp < 1 || model->ncomp > kMaxComp) return set_error(ctx, SS_ERR_INVALID, "ncomp must be in 1..%d", kMaxComp);
    if (model->ncoeffs < 1 || model->ncoeffs > SS_MAX_NCOEFFS) return set_error(ctx, SS_ERR_INVALID, "model ncoeffs must be in 1..%d", SS_MAX_NCOEFFS);
    if (!model->means || !model->covs || !model->weights) return set_error(ctx, SS_ERR_INVALID, "model arrays are NULL");
    return SS_OK;
}

// d_sym: n symbols on the device. Leaves votes in st->d_votes (n+1), lens in st->d_lens; returns nseg.
static int vote_split_dev(ss_ctx* ctx, SegState* st, const uint8_t* d_sym, size_t n, int depth, int threshold, size_t* out_nseg) {
    SS_CUDA(ctx, st->d_votes.reserve(n + 1));
    SS_CUDA(ctx, cudaMemsetAsync(st->d_votes.p, 0, (n + 1) * sizeof(uint32_t), ctx->stream));
    *out_nseg = 0;
    if (n == 0) return SS_OK;
    const int maxlen = depth + 1;
    if (n >= (size_t)depth) {
        SS_CUDA(ctx, st->d_keys.reserve(n));
        SS_CUDA(ctx, st->d_keys_sorted.reserve(n));
        SS_CUDA(ctx, st->d_nruns.reserve(kMaxDepth + 2));
        SS_CUDA(ctx, cudaMemsetAsync(st->d_nruns.p, 0, (kMaxDepth + 2) * sizeof(uint32_t), ctx->stream));
        SS_CUDA(ctx, st->d_isums.reserve(2 * (kMaxDepth + 2)));
        SS_CUDA(ctx, cudaMemsetAsync(st->d_isums.p, 0, 2 * (kMaxDepth + 2) * sizeof(unsigned long long), ctx->stream));
        SS_CUDA(ctx, st->d_h.reserve(n));
        SS_CUDA(ctx, st->d_stats.reserve(48));
        const int tb = 256;
        for (int len = 1; len <= maxlen; len++) {
            if ((size_t)len > n) break;
            const size_t cntk = n - len + 1;
            SS_CUDA(ctx, st->d_uniq[len].reserve(cntk));
            SS_CUDA(ctx, st->d_cnt[len].reserve(cntk));
            k_ve_keys<<<ceil_div((long long)cntk, tb), tb, 0, ctx->stream>>>(d_sym, n, len, st->d_keys.p);
            SS_LAUNCHED(ctx);
            size_t tmp1 = 0, tmp2 = 0;
            cub::DeviceRadixSort::SortKeys(nullptr, tmp1, st->d_keys.p, st->d_keys_sorted.p, (int)cntk, 0, 8 * len, ctx->stream);
            cub::DeviceRunLengthEncode::Encode(nullptr, tmp2, st->d_keys_sorted.p, st->d_uniq[len].p, st->d_cnt[len].p,
                                               st->d_nruns.p + len, (int)cntk, ctx->stream);
            size_t tmp = std::max(tmp1, tmp2);
            SS_CUDA(ctx, st->d_cub_tmp.reserve(tmp));
            SS_CUDA(ctx, cub::DeviceRadixSort::SortKeys(st->d_cub_tmp.p, tmp, st->d_keys.p, st->d_keys_sorted.p, (int)cntk, 0, 8 * len,
                                                        ctx->stream));
            ctx->launches++;
            SS_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(st->d_cub_tmp.p, tmp, st->d_keys_sorted.p, st->d_uniq[len].p,
                                                            st->d_cnt[len].p, st->d_nruns.p + len, (int)cntk, ctx->stream));
            ctx->launches++;
        }
        const int g = ctx->sm_count * 4;
        for (int len = 1; len <= depth; len++) {
            const size_t cntk = n - len + 1;
            SS_CUDA(ctx, st->d_zf[len].reserve(cntk));
            SS_CUDA(ctx, st->d_zh[len].reserve(cntk));
            SS_CUDA(ctx, st->d_rank[len].reserve(cntk));
            k_ve_isums<<<g, tb, 0, ctx->stream>>>(st->d_cnt[len].p, st->d_nruns.p + len, st->d_isums.p + 2 * len);
            SS_LAUNCHED(ctx);
            k_ve_zfreq<<<g, tb, 0, ctx->stream>>>(st->d_cnt[len].p, st->d_nruns.p + len, st->d_isums.p + 2 * len, st->d_zf[len].p);
            SS_LAUNCHED(ctx);
            const bool has_next = (size_t)(len + 1) <= n;
            k_ve_entropy<<<g, tb, 0, ctx->stream>>>(st->d_uniq[len].p, st->d_nruns.p + len, has_next ? st->d_uniq[len + 1].p : nullptr,
                                                   has_next ? st->d_cnt[len + 1].p : nullptr, has_next ? st->d_nruns.p + len + 1 : nullptr,
                                                   st->d_h.p);
            SS_LAUNCHED(ctx);
            k_ve_hstats<<<1, 256, 0, ctx->stream>>>(st->d_h.p, st->d_nruns.p + len, st->d_stats.p + 40);
            SS_LAUNCHED(ctx);
            k_ve_zent<<<g, tb, 0, ctx->stream>>>(st->d_h.p, st->d_nruns.p + len, st->d_stats.p + 40, st->d_zh[len].p);
            SS_LAUNCHED(ctx);
            k_ve_rank<<<ceil_div((long long)cntk, tb), tb, 0, ctx->stream>>>(d_sym, n, len, st->d_uniq[len].p, st->d_nruns.p + len,
                                                                           st->d_rank[len].p);
            SS_LAUNCHED(ctx);
        }
        VeTables t;
        for (int len = 0; len < kMaxDepth + 2; len++) {
            t.zf[len] = st->d_zf[len].p;
            t.zh[len] = st->d_zh[len].p;
            t.rank[len] = st->d_rank[len].p;
        }
        const size_t nwin = n - depth + 1;
        k_ve_votes<<<ceil_div((long long)nwin, tb), tb, 0, ctx->stream>>>(n, depth, t, st->d_votes.p);
        SS_LAUNCHED(ctx);
    }
    // split_string
    SS_CUDA(ctx, st->d_flags.reserve(n));
    SS_CUDA(ctx, st->d_bounds.reserve(n));
    SS_CUDA(ctx, st->d_nbounds.reserve(1));
    SS_CUDA(ctx, st->d_lens.reserve(n));
    k_ve_flags<<<ceil_div((long long)n, 256), 256, 0, ctx->stream>>>(st->d_votes.p, n, (uint32_t)std::max(threshold, 0), st->d_flags.p);
    SS_LAUNCHED(ctx);
    cub::CountingInputIterator<uint32_t> counting(0);
    size_t tmp = 0;
    cub::DeviceSelect::Flagged(nullptr, tmp, counting, st->d_flags.p, st->d_bounds.p, st->d_nbounds.p, (int)n, ctx->stream);
    SS_CUDA(ctx, st->d_cub_tmp.reserve(tmp));
    SS_CUDA(ctx, cub::DeviceSelect::Flagged(st->d_cub_tmp.p, tmp, counting, st->d_flags.p, st->d_bounds.p, st->d_nbounds.p, (int)n,
                                            ctx->stream));
    ctx->launches++;
    k_ve_lens<<<ctx->sm_count, 256, 0, ctx->stream>>>(st->d_bounds.p, st->d_nbounds.p, n, (uint64_t)SS_HOP, st->d_lens.p);
    SS_LAUNCHED(ctx);
    uint32_t nb = 0;
    SS_CUDA(ctx, cudaMemcpyAsync(&nb, st->d_nbounds.p, sizeof(nb), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out_nseg = (size_t)nb + 1;
    return SS_OK;
}

static int check_ve(ss_ctx* ctx, int depth, size_t n) {
    if (depth < 1 || depth > kMaxDepth) return set_error(ctx, SS_ERR_INVALID, "depth must be in 1..%d (got %d)", kMaxDepth, depth);
    if (n > 0x7FFFFFF0ull) return set_error(ctx, SS_ERR_INVALID, "too many symbols");
    return SS_OK;
}

}  // namespace ss

using namespace ss;

extern "C" {

int ss_symbols(ss_ctx* ctx, const double* mfcc, size_t frames, const ss_gmm* model, uint8_t* out_symbols, double* out_posteriors) {
    if (!ctx) return set_error(nullptr, SS_ERR_INVALID, "ctx is NULL");
    SS_TRY(check_model(ctx, model));
    if (frames < 2) return set_error(ctx, SS_ERR_TOO_FEW_ROWS, "Standardizer needs at least 2 rows (got %zu)", frames);
    if (!mfcc || !out_symbols) return set_error(ctx, SS_ERR_INVALID, "NULL buffer");
    SS_CUDA(ctx, cudaSetDevice(ctx->device));
    SegState* st = seg_state(ctx);
    const int c = model->ncoeffs;
    SS_TRY(upload(ctx, st->d_mfcc, mfcc, frames * (size_t)c));
    SS_CUDA(ctx, st->d_sym.reserve(frames));
    if (out_posteriors) SS_CUDA(ctx, st->d_post.reserve(frames * (size_t)model->ncomp));
    SS_TRY(symbols_dev(ctx, st, st->d_mfcc.p, frames, model, st->d_sym.p, out_posteriors ? st->d_post.p : nullptr));
    SS_CUDA(ctx, cudaMemcpyAsync(out_symbols, st->d_sym.p, frames, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_posteriors)
        SS_CUDA(ctx, cudaMemcpyAsync(out_posteriors, st->d_post.p, frames * (size_t)model->ncomp * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int ss_vote_split(ss_ctx* ctx, const uint8_t* symbols, size_t n, int depth, int threshold, uint32_t* out_votes, uint64_t* out_seg_lens,
                  size_t* out_nseg) {
    if (!ctx) return set_error(nullptr, SS_ERR_INVALID, "ctx is NULL");
    SS_TRY(check_ve(ctx, depth, n));
    if (!out_nseg || (n && (!symbols || !out_seg_lens))) return set_error(ctx, SS_ERR_INVALID, "NULL buffer");
    SS_CUDA(ctx, cudaSetDevice(ctx->device));
    SegState* st = seg_state(ctx);
    SS_TRY(upload(ctx, st->d_sym, symbols, n));
    size_t nseg = 0;
    SS_TRY(vote_split_dev(ctx, st, st->d_sym.p, n, depth, threshold, &nseg));
    if (out_votes) SS_CUDA(ctx, cudaMemcpyAsync(out_votes, st->d_votes.p, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (nseg) SS_CUDA(ctx, cudaMemcpyAsync(out_seg_lens, st->d_lens.p, nseg * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out_nseg = nseg;
    return SS_OK;
}

int ss_partition(ss_ctx* ctx, const double* mfcc, size_t frames, const ss_gmm* model, int depth, int threshold, uint64_t* out_seg_lens,
                 size_t* out_nseg) {
    if (!ctx) return set_error(nullptr, SS_ERR_INVALID, "ctx is NULL");
    SS_TRY(check_model(ctx, model));
    SS_TRY(check_ve(ctx, depth, frames));
    if (frames < 2) return set_error(ctx, SS_ERR_TOO_FEW_ROWS, "Standardizer needs at least 2 rows (got %zu)", frames);
    if (!mfcc || !out_seg_lens || !out_nseg) return set_error(ctx, SS_ERR_INVALID, "NULL buffer");
    SS_CUDA(ctx, cudaSetDevice(ctx->device));
    SegState* st = seg_state(ctx);
    SS_TRY(upload(ctx, st->d_mfcc, mfcc, frames * (size_t)model->ncoeffs));
    SS_CUDA(ctx, st->d_sym.reserve(frames));
    SS_TRY(symbols_dev(ctx, st, st->d_mfcc.p, frames, model, st->d_sym.p, nullptr));  // symbols stay in HBM
    size_t nseg = 0;
    SS_TRY(vote_split_dev(ctx, st, st->d_sym.p, frames, depth, threshold, &nseg));
    if (nseg) SS_CUDA(ctx, cudaMemcpyAsync(out_seg_lens, st->d_lens.p, nseg * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out_nseg = nseg;
    return SS_OK;
}

int ss_sequence_distances(ss_ctx* ctx, const double* mean_mfccs, size_t nrows, int ncoeffs, double* out_dist) {
    if (!ctx) return set_error(nullptr, SS_ERR_INVALID, "ctx is NULL");
    if (ncoeffs < 1 || ncoeffs > SS_MAX_NCOEFFS) return set_error(ctx, SS_ERR_INVALID, "ncoeffs out of range");
    if (nrows < 2) return SS_OK;
    if (!mean_mfccs || !out_dist) return set_error(ctx, SS_ERR_INVALID, "NULL buffer");
    SS_CUDA(ctx, cudaSetDevice(ctx->device));
    SegState* st = seg_state(ctx);
    SS_TRY(upload(ctx, st->d_mfcc, mean_mfccs, nrows * (size_t)ncoeffs));
    SS_CUDA(ctx, st->d_z.reserve(nrows));
    k_seq_dist<<<ceil_div((long long)nrows, 128), 128, 0, ctx->stream>>>(st->d_mfcc.p, nrows, ncoeffs, st->d_z.p);
    SS_LAUNCHED(ctx);
    SS_CUDA(ctx, cudaMemcpyAsync(out_dist, st->d_z.p, (nrows - 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int ss_gmm_train(ss_ctx* ctx, const double* mfcc, size_t frames, int ncoeffs, int ncomp, int iters, double reg, uint64_t seed,
                 double* out_means, double* out_covs, double* out_weights) {
    if (!ctx) return set_error(nullptr, SS_ERR_INVALID, "ctx is NULL");
    if (ncoeffs < 1 || ncoeffs > SS_MAX_NCOEFFS) return set_error(ctx, SS_ERR_INVALID, "ncoeffs must be in 1..%d", SS_MAX_NCOEFFS);
    if (ncomp < 1 || ncomp > kMaxComp) return set_error(ctx, SS_ERR_INVALID, "ncomp must be in 1..%d", kMaxComp);
    if (iters < 0) return set_error(ctx, SS_ERR_INVALID, "iters must be >= 0");
    if (frames < 2 || frames < (size_t)ncomp) return set_error(ctx, SS_ERR_TOO_FEW_ROWS, "training needs >= max(2, ncomp) rows (got %zu)", frames);
    if (!mfcc || !out_means || !out_covs || !out_weights) return set_error(ctx, SS_ERR_INVALID, "NULL buffer");
    SS_CUDA(ctx, cudaSetDevice(ctx->device));
    SegState* st = seg_state(ctx);
    const int c = ncoeffs, nout = c * c + c + 1;
    SS_TRY(upload(ctx, st->d_mfcc, mfcc, frames * (size_t)c));
    // Standardizer (A5)
    const int nb = (int)std::min<size_t>(std::max<size_t>(frames / 64, 1), 1024);
    SS_CUDA(ctx, st->d_partial.reserve(std::max<size_t>((size_t)nb * 16, (size_t)ncomp * 256 * nout)));
    SS_CUDA(ctx, st->d_stats.reserve(48));
    double* d_mean = st->d_stats.p;
    double* d_var = d_mean + 16;
    double* d_sd = d_mean + 32;
    k_col_partial<<<nb, 256, 0, ctx->stream>>>(st->d_mfcc.p, frames, c, nullptr, 0, st->d_partial.p);
    SS_LAUNCHED(ctx);
    k_col_final<<<1, 32, 0, ctx->stream>>>(st->d_partial.p, nb, c, (double)frames, d_mean, nullptr);
    SS_LAUNCHED(ctx);
    k_col_partial<<<nb, 256, 0, ctx->stream>>>(st->d_mfcc.p, frames, c, d_mean, 1, st->d_partial.p);
    SS_LAUNCHED(ctx);
    k_col_final<<<1, 32, 0, ctx->stream>>>(st->d_partial.p, nb, c, (double)(frames - 1), d_var, d_sd);
    SS_LAUNCHED(ctx);
    SS_CUDA(ctx, st->d_z.reserve(frames * (size_t)c));
    k_standardize<<<ceil_div((long long)(frames * c), 256), 256, 0, ctx->stream>>>(st->d_mfcc.p, frames, c, d_mean, d_sd, st->d_z.p);
    SS_LAUNCHED(ctx);
    // model buffers: means | covs | weights, moments, inverse covariances
    const size_t msz = (size_t)ncomp * c + (size_t)ncomp * c * c + ncomp;
    SS_CUDA(ctx, st->d_model.reserve(msz));
    double* d_means = st->d_model.p;
    double* d_covs = d_means + (size_t)ncomp * c;
    double* d_w = d_covs + (size_t)ncomp * c * c;
    SS_CUDA(ctx, st->d_inv.reserve((size_t)ncomp * c * c));
    SS_CUDA(ctx, st->d_sqrt_det.reserve(ncomp));
    SS_CUDA(ctx, st->d_status.reserve(1));
    SS_CUDA(ctx, cudaMemsetAsync(st->d_status.p, 0, sizeof(int), ctx->stream));
    SS_CUDA(ctx, st->d_post.reserve(frames * (size_t)ncomp));
    SS_CUDA(ctx, st->d_sym.reserve(frames));
    SS_CUDA(ctx, st->d_h.reserve((size_t)ncomp * nout + 64));  // moments
    double* d_mom = st->d_h.p;
    const int mb = (int)std::min<size_t>((frames + 31) / 32, 256);
    // initial covariance: column means of z, then centred second moments (mode 2)
    k_col_partial<<<nb, 256, 0, ctx->stream>>>(st->d_z.p, frames, c, nullptr, 0, st->d_partial.p);
    SS_LAUNCHED(ctx);
    k_col_final<<<1, 32, 0, ctx->stream>>>(st->d_partial.p, nb, c, (double)frames, d_var, nullptr);  // d_var reused: mean of z
    SS_LAUNCHED(ctx);
    k_gmm_moments<<<dim3(mb, 1), 256, 0, ctx->stream>>>(st->d_z.p, frames, c, ncomp, nullptr, d_var, 2, st->d_partial.p);
    SS_LAUNCHED(ctx);
    k_gmm_moments_final<<<1, 256, 0, ctx->stream>>>(st->d_partial.p, mb, nout, d_mom);
    SS_LAUNCHED(ctx);
    k_gmm_init_covs<<<ncomp, 256, 0, ctx->stream>>>(d_mom, c, ncomp, frames, reg, d_covs);
    SS_LAUNCHED(ctx);
    // initial means: `ncomp` distinct rows, partial Fisher-Yates with a seeded mt19937_64 (the reference uses thread_rng)
    std::vector<uint64_t> pick(ncomp);
    {
        std::mt19937_64 rng(seed);
        std::vector<size_t> idx(frames);
        for (size_t i = 0; i < frames; i++) idx[i] = i;
        for (int j = 0; j < ncomp; j++) {
            std::uniform_int_distribution<size_t> dist(j, frames - 1);
            std::swap(idx[j], idx[dist(rng)]);
            pick[j] = idx[j];
        }
    }
    SS_TRY(upload(ctx, st->d_lens, pick.data(), (size_t)ncomp));
    k_gmm_init_means<<<ncomp, 32, 0, ctx->stream>>>(st->d_z.p, c, st->d_lens.p, d_means, d_w, ncomp);
    SS_LAUNCHED(ctx);
    const size_t smem = sizeof(double) * ((size_t)ncomp * c * c + (size_t)ncomp * c + 2 * (size_t)ncomp);
    SS_CUDA(ctx, cudaFuncSetAttribute(k_gmm_symbols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int it = 0; it < iters; it++) {
        k_gmm_prepare<<<1, 32, 0, ctx->stream>>>(d_covs, ncomp, c, st->d_inv.p, st->d_sqrt_det.p, st->d_status.p);
        SS_LAUNCHED(ctx);
        k_gmm_symbols<<<ceil_div((long long)frames, 128), 128, smem, ctx->stream>>>(st->d_z.p, frames, c, ncomp, d_means, st->d_inv.p,
                                                                                  st->d_sqrt_det.p, d_w, st->d_sym.p, st->d_post.p);
        SS_LAUNCHED(ctx);
        k_gmm_moments<<<dim3(mb, ncomp), 256, 0, ctx->stream>>>(st->d_z.p, frames, c, ncomp, st->d_post.p, nullptr, 0, st->d_partial.p);
        SS_LAUNCHED(ctx);
        k_gmm_moments_final<<<ncomp, 256, 0, ctx->stream>>>(st->d_partial.p, mb, nout, d_mom);
        SS_LAUNCHED(ctx);
        k_gmm_update_means<<<ncomp, 32, 0, ctx->stream>>>(d_mom, c, ncomp, frames, d_means, d_w, st->d_status.p);
        SS_LAUNCHED(ctx);
        k_gmm_moments<<<dim3(mb, ncomp), 256, 0, ctx->stream>>>(st->d_z.p, frames, c, ncomp, st->d_post.p, d_means, 1, st->d_partial.p);
        SS_LAUNCHED(ctx);
        k_gmm_moments_final<<<ncomp, 256, 0, ctx->stream>>>(st->d_partial.p, mb, nout, d_mom);
        SS_LAUNCHED(ctx);
        k_gmm_update_covs<<<ncomp, 256, 0, ctx->stream>>>(d_mom, c, reg, d_covs);
        SS_LAUNCHED(ctx);
    }
    k_gmm_prepare<<<1, 32, 0, ctx->stream>>>(d_covs, ncomp, c, st->d_inv.p, st->d_sqrt_det.p, st->d_status.p);  // predict must work
    SS_LAUNCHED(ctx);
    k_gmm_check_det<<<1, 32, 0, ctx->stream>>>(st->d_sqrt_det.p, ncomp, st->d_status.p);
    SS_LAUNCHED(ctx);
    int status = 0;
    SS_CUDA(ctx, cudaMemcpyAsync(&status, st->d_status.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaMemcpyAsync(out_means, d_means, sizeof(double) * ncomp * c, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaMemcpyAsync(out_covs, d_covs, sizeof(double) * ncomp * c * c, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaMemcpyAsync(out_weights, d_w, sizeof(double) * ncomp, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (status != 0)  // the reference retries with a fresh random draw (`while let Err`, src/lib.rs:50-52); the caller does the same with seed + 1
        return set_error(ctx, SS_ERR_INVALID, "EM failed (status %d): a component collapsed or a covariance became singular", status);
    return SS_OK;
}

}  // extern "C"
