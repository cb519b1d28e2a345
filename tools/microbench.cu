// Microbenchmarks that settle the instruction-mix questions behind the DTW kernel design
// (DESIGN.md "DTW kernel"): FP32 FMA issue rate scalar vs packed f32x2, broadcast LDS rate by width,
// three-input min. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float s) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 0.001f + i;
    float a0 = s, a1 = s * 1.1f, a2 = s * 1.2f, a3 = s * 1.3f;
    float b0 = s + 1.f, b1 = s + 2.f, b2 = s + 3.f, b3 = s + 4.f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            acc[i] = fmaf(a0, b0, acc[i]);
            acc[i + 1] = fmaf(a1, b1, acc[i + 1]);
            acc[i + 2] = fmaf(a2, b2, acc[i + 2]);
            acc[i + 3] = fmaf(a3, b3, acc[i + 3]);
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
    return ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(x);
}

__global__ void k_ffma2(float* out, float s) {
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = pack2(threadIdx.x * 0.001f + i, i);
    unsigned long long a0 = pack2(s, s * 1.1f), a1 = pack2(s * 1.2f, s * 1.3f);
    unsigned long long b0 = pack2(s + 1.f, s + 2.f), b1 = pack2(s + 3.f, s + 4.f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            acc[i] = fma2(a0, b0, acc[i]);
            acc[i + 1] = fma2(a1, b1, acc[i + 1]);
        }
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            acc[i] = fma2(a0, b0, acc[i]);
            acc[i + 1] = fma2(a1, b1, acc[i + 1]);
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// broadcast LDS: every lane reads the same address. W = 1,2,4 words.
template <int W>
__global__ void k_lds_bcast(float* out, int stride) {
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 0.5f;
    __syncthreads();
    float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    int idx = 0;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            int o = (idx + u * 4 * stride) & 4095 & ~3;
            if (W == 1) { acc0 += sm[o]; }
            if (W == 2) { float2 v = *reinterpret_cast<float2*>(&sm[o]); acc0 += v.x; acc1 += v.y; }
            if (W == 4) { float4 v = *reinterpret_cast<float4*>(&sm[o]); acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w; }
        }
        idx = (idx + 32 * stride) & 4095;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

// per-lane distinct LDS.128 (conflict-free 16B per lane contiguous)
__global__ void k_lds_lane128(float* out) {
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 0.5f;
    __syncthreads();
    float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    int idx = (threadIdx.x & 31) * 4;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            float4 v = *reinterpret_cast<float4*>(&sm[(idx + u * 128) & 4095]);
            acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w;
        }
        idx = (idx + 1024) & 4095;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

// min3 + add chain mix: 8 independent chains
__global__ void k_min3(float* out, float s) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x + i * s;
    float p = s, q = s * 2;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float m;
            asm("min.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(v[i]), "f"(p), "f"(q));
            v[i] = m + s;
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// the DTW inner pattern, k-paired: per cell 3 LDS.128 + 1 LDS.64 broadcast, 7 FFMA2, 1 FADD, min3, add
template <int R>
__global__ void k_dtw_pattern(float* out, float s) {
    __shared__ __align__(16) float sm[32 * 16];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = i * 0.01f;
    __syncthreads();
    unsigned long long a[R][7];
    float dprev[R][32];
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int k = 0; k < 7; k++) a[r][k] = pack2(s * k + threadIdx.x * 0.01f + r, s + k);
#pragma unroll
        for (int j = 0; j < 32; j++) dprev[r][j] = 1e30f;
    }
    for (int it = 0; it < ITERS / 32; it++) {
        float left[R], diag[R];
#pragma unroll
        for (int r = 0; r < R; r++) { left[r] = 1e30f; diag[r] = it ? 1e30f : 0.f; }
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const float4* bp = reinterpret_cast<const float4*>(&sm[j * 16]);
            float4 b0 = bp[0], b1 = bp[1], b2 = bp[2];
            float2 b3 = *reinterpret_cast<const float2*>(&sm[j * 16 + 12]);
            unsigned long long bb[7] = {pack2(b0.x, b0.y), pack2(b0.z, b0.w), pack2(b1.x, b1.y), pack2(b1.z, b1.w),
                                        pack2(b2.x, b2.y), pack2(b2.z, b2.w), pack2(b3.x, b3.y)};
#pragma unroll
            for (int r = 0; r < R; r++) {
                unsigned long long acc = pack2(s, 0.f);
#pragma unroll
                for (int k = 0; k < 7; k++) acc = fma2(a[r][k], bb[k], acc);
                float c = __uint_as_float((unsigned)acc) + __uint_as_float((unsigned)(acc >> 32));
                float up = dprev[r][j];
                float m;
                asm("min.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(left[r]), "f"(up), "f"(diag[r]));
                float cur = c + m;
                diag[r] = up; dprev[r][j] = cur; left[r] = cur;
            }
        }
    }
    float r0 = 0;
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < 32; j++) r0 += dprev[r][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r0;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("device %s sms %d clock %d kHz\n", p.name, sms, khz);
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024));
    for (int tpb : {128, 256, 512}) {
        for (int bps : {1, 2, 4}) {
            if (tpb * bps > 2048) continue;
            int grid = sms * bps;
            double lanes = (double)grid * tpb;
            float ms;
            ms = timeit([&] { k_ffma<<<grid, tpb>>>(out, 1.0001f); });
            printf("tpb %d bps %d  ffma   : %.3f ms  %.2f Tfma/s  (%.1f fma/clk/SM @1.965GHz)\n", tpb, bps, ms, lanes * ITERS * 16 / ms / 1e9, lanes * ITERS * 16 / (ms * 1e-3) / sms / 1.965e9);
            ms = timeit([&] { k_ffma2<<<grid, tpb>>>(out, 1.0001f); });
            printf("tpb %d bps %d  ffma2  : %.3f ms  %.2f Tfma/s  (%.1f fma/clk/SM)\n", tpb, bps, ms, lanes * ITERS * 32 / ms / 1e9, lanes * ITERS * 32 / (ms * 1e-3) / sms / 1.965e9);
            ms = timeit([&] { k_lds_bcast<1><<<grid, tpb>>>(out, 1); });
            printf("tpb %d bps %d  lds32 b: %.3f ms  %.2f warp-LDS/clk/SM\n", tpb, bps, ms, lanes / 32 * ITERS * 8 / (ms * 1e-3) / sms / 1.965e9);
            ms = timeit([&] { k_lds_bcast<2><<<grid, tpb>>>(out, 1); });
            printf("tpb %d bps %d  lds64 b: %.3f ms  %.2f warp-LDS/clk/SM\n", tpb, bps, ms, lanes / 32 * ITERS * 8 / (ms * 1e-3) / sms / 1.965e9);
            ms = timeit([&] { k_lds_bcast<4><<<grid, tpb>>>(out, 1); });
            printf("tpb %d bps %d  lds128b: %.3f ms  %.2f warp-LDS/clk/SM\n", tpb, bps, ms, lanes / 32 * ITERS * 8 / (ms * 1e-3) / sms / 1.965e9);
            ms = timeit([&] { k_lds_lane128<<<grid, tpb>>>(out); });
            printf("tpb %d bps %d  lds128l: %.3f ms  %.2f warp-LDS/clk/SM\n", tpb, bps, ms, lanes / 32 * ITERS * 8 / (ms * 1e-3) / sms / 1.965e9);
            ms = timeit([&] { k_min3<<<grid, tpb>>>(out, 1.0001f); });
            printf("tpb %d bps %d  min3+add: %.3f ms  %.2f pairs/clk/SM\n", tpb, bps, ms, lanes * ITERS * 8 / (ms * 1e-3) / sms / 1.965e9);
            ms = timeit([&] { k_dtw_pattern<1><<<grid, tpb>>>(out, 1.0001f); });
            printf("tpb %d bps %d  dtwpat R1: %.3f ms  %.3e cells/s (%.2f cells/clk/SM)\n", tpb, bps, ms, lanes * ITERS / (ms * 1e-3), lanes * ITERS / (ms * 1e-3) / sms / 1.965e9);
            if (tpb <= 256) {
                ms = timeit([&] { k_dtw_pattern<2><<<grid, tpb>>>(out, 1.0001f); });
                printf("tpb %d bps %d  dtwpat R2: %.3f ms  %.3e cells/s (%.2f cells/clk/SM)\n", tpb, bps, ms, lanes * ITERS * 2 / (ms * 1e-3), lanes * ITERS * 2 / (ms * 1e-3) / sms / 1.965e9);
            }
        }
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
