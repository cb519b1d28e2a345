// How does tcgen05.mma kind::f16 behave with an F16 accumulator (instruction-descriptor c_format = 0), and where do the
// 16-bit results sit in tensor memory? One CTA, one M=128 x N=128 x K=16 MMA on operands shaped like the DTW scan's
// (A = [-2(a-mu) (13), s, s, rd(|a|^2/s)], B = [b-mu (13), hi, lo, s]), once into an F32 accumulator and once into an F16
// one; TMEM is read back raw (32x32b .b32) and with .pack::16b, and the host compares with the f64 sum of the exact products.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_f16acc microbench_f16acc.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__host__ __device__ inline int tile_offset(int row, int k) { return ((k >> 3) * 16 + (row >> 3)) * 128 + (row & 7) * 16 + (k & 7) * 2; }
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k_probe(const unsigned char* __restrict__ gA, const unsigned char* __restrict__ gB, uint32_t* __restrict__ raw16,
                                                  uint32_t* __restrict__ packed16, float* __restrict__ f32out, int variant) {
    __shared__ __align__(128) unsigned char sA[4096];
    __shared__ __align__(128) unsigned char sB[4096];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += 128) {
        reinterpret_cast<uint32_t*>(sA)[i] = reinterpret_cast<const uint32_t*>(gA)[i];
        reinterpret_cast<uint32_t*>(sB)[i] = reinterpret_cast<const uint32_t*>(gB)[i];
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        auto desc = [](uint32_t saddr) { return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((128 / 8 * 128) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46); };
        const uint64_t ad = desc(s32(sA)), bd = desc(s32(sB));
        const uint32_t idesc32 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // D = F32
        const uint32_t idesc16 = (0u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // D = F16
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + 128), "l"(ad), "l"(bd), "r"(idesc32), "r"(0u) : "memory");
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc16), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < 128; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(lane_addr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) raw16[m * 128 + c0 + j] = r[j];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(lane_addr + 128 + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) f32out[m * 128 + c0 + j] = __uint_as_float(r[j]);
    }
    // .pack::16b: x8 registers <- 16 columns
    for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(lane_addr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) packed16[m * 64 + c0 / 2 + j] = r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }
static uint16_t f2h(float f) { __half x = __float2half_rn(f); uint16_t h; memcpy(&h, &x, 2); return h; }

int main() {
    std::mt19937_64 rng(7);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<uint16_t> A(128 * 16), B(128 * 16);
    const float scale = 4.f;
    const double sig[13] = {20, 10, 6.7, 5, 4, 3.3, 2.9, 2.5, 2.2, 2, 1.8, 1.7, 1.5};
    for (int m = 0; m < 128; m++) {
        double na = 0, nb = 0;
        const double amp = (m % 8 == 7) ? 4.0 : 1.4;  // a few loud rows
        for (int k = 0; k < 13; k++) {
            const float a = h2f(f2h((float)(nd(rng) * sig[k] * amp))), b = h2f(f2h((float)(nd(rng) * sig[k] * amp)));
            na += (double)a * a, nb += (double)b * b;
            A[m * 16 + k] = f2h(-2.f * a);
            B[m * 16 + k] = f2h(b);
        }
        __half rd = __float2half_rd((float)na / scale);
        memcpy(&A[m * 16 + 15], &rd, 2);
        A[m * 16 + 13] = A[m * 16 + 14] = f2h(scale);
        const float sn = (float)nb / scale;
        B[m * 16 + 13] = f2h(sn);
        B[m * 16 + 14] = f2h(sn - h2f(f2h(sn)));
        B[m * 16 + 15] = f2h(scale);
    }
    std::vector<unsigned char> tA(4096), tB(4096);
    for (int r = 0; r < 128; r++)
        for (int k = 0; k < 16; k++) {
            memcpy(&tA[tile_offset(r, k)], &A[r * 16 + k], 2);
            memcpy(&tB[tile_offset(r, k)], &B[r * 16 + k], 2);
        }
    unsigned char *dA, *dB;
    uint32_t *draw, *dpk;
    float* df32;
    CK(cudaMalloc(&dA, 4096));
    CK(cudaMalloc(&dB, 4096));
    CK(cudaMalloc(&draw, 128 * 128 * 4));
    CK(cudaMalloc(&dpk, 128 * 64 * 4));
    CK(cudaMalloc(&df32, 128 * 128 * 4));
    CK(cudaMemcpy(dA, tA.data(), 4096, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, tB.data(), 4096, cudaMemcpyHostToDevice));
    k_probe<<<1, 128>>>(dA, dB, draw, dpk, df32, 0);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> raw(128 * 128), pk(128 * 64);
    std::vector<float> f32(128 * 128);
    CK(cudaMemcpy(raw.data(), draw, raw.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pk.data(), dpk, pk.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(f32.data(), df32, f32.size() * 4, cudaMemcpyDeviceToHost));
    printf("raw words of the F16 accumulator, lane 0, columns 0..7: ");
    for (int j = 0; j < 8; j++) printf("%08x ", raw[j]);
    printf("\npacked words, lane 0, registers 0..3:                     ");
    for (int j = 0; j < 4; j++) printf("%08x ", pk[j]);
    printf("\n");
    long hi_nonzero = 0, pack_ok = 0, eq_rn_exact = 0, eq_rn_f32 = 0, eq_rz_exact = 0, total = 0;
    double max_rel16 = 0, max_rel32 = 0, max_abs32_over_norm = 0, max_ulp = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < 128; n++) {
            double exact = 0, mag = 0;
            for (int k = 0; k < 16; k++) {
                const double p = (double)h2f(A[m * 16 + k]) * (double)h2f(B[n * 16 + k]);
                exact += p;
                mag += fabs(p);
            }
            const uint32_t w = raw[m * 128 + n];
            hi_nonzero += (w >> 16) != 0;
            const uint16_t h = (uint16_t)(w & 0xFFFF);
            const uint32_t pw = pk[m * 64 + n / 2];
            pack_ok += (uint16_t)((n & 1) ? (pw >> 16) : (pw & 0xFFFF)) == h;
            const float v16 = h2f(h), v32 = f32[m * 128 + n];
            eq_rn_exact += h == f2h((float)exact);
            eq_rn_f32 += h == f2h(v32);
            __half rz = __float2half_rz((float)exact);
            uint16_t hz;
            memcpy(&hz, &rz, 2);
            eq_rz_exact += h == hz;
            const double ulp = ldexp(1.0, (int)floor(log2(fmax(fabs(exact), 6.2e-5))) - 10);
            max_ulp = fmax(max_ulp, fabs(v16 - exact) / ulp);
            max_rel16 = fmax(max_rel16, fabs(v16 - exact) / fmax(fabs(exact), 1e-3));
            max_rel32 = fmax(max_rel32, fabs(v32 - exact) / fmax(fabs(exact), 1e-3));
            max_abs32_over_norm = fmax(max_abs32_over_norm, fabs(v32 - exact) / mag);
            total++;
        }
    printf("pairs %ld: high half of the raw word non-zero in %ld; pack::16b register halves equal the raw low halves in %ld\n", total, hi_nonzero, pack_ok);
    printf("F16 accumulator == rn_f16(exact sum) in %ld, == rn_f16(F32 accumulator) in %ld, == rz_f16(exact) in %ld of %ld\n", eq_rn_exact, eq_rn_f32,
           eq_rz_exact, total);
    printf("F16 accumulator: max |err| = %.3f ulp_f16(exact), max relative %.3e;  F32 accumulator: max relative %.3e, max |err| / sum|products| = %.3e\n",
           max_ulp, max_rel16, max_rel32, max_abs32_over_norm);
    return 0;
}
