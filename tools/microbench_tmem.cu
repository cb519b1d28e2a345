// TMEM read throughput (tcgen05.ld 32x32b.x32 / .x16 / .x8) for W warps per SM: bytes / clk / SM. Each warp reads its
// own lane quadrant (warp % 4). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_tmem microbench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int W>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[32]);
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
template <>
__device__ __forceinline__ void ld<8>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

template <int W, int PER_WAIT>
__global__ void __launch_bounds__(512, 1) k_tmem(uint32_t* out, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < PER_WAIT; k++) {
            ld<W>(base + (((it + k) * W) & 255), r);
            if (k == PER_WAIT - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += r[0] ^ r[W - 1];
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int W, int PER_WAIT>
static int run(int warps, int sms, uint32_t* out) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_tmem<W, PER_WAIT><<<sms, warps * 32>>>(out, iters);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_tmem<W, PER_WAIT><<<sms, warps * 32>>>(out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)warps * 32 * W * 4 * iters * PER_WAIT;  // per SM
    printf("x%-2d per-wait %d warps/SM %2d: %.3f ms  %.1f B/clk/SM (%.1f per SMSP)\n", W, PER_WAIT, warps, ms, bytes / (ms * 1e-3) / 1.965e9,
           bytes / (ms * 1e-3) / 1.965e9 / (warps < 4 ? warps : 4));
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t* out;
    CK(cudaMalloc(&out, (size_t)sms * 512 * 4));
    for (int warps : {1, 4, 8, 12, 16}) {
        if (run<32, 1>(warps, sms, out)) return 1;
        if (run<32, 2>(warps, sms, out)) return 1;
        if (run<16, 2>(warps, sms, out)) return 1;
        if (run<8, 4>(warps, sms, out)) return 1;
    }
    printf("done\n");
    return 0;
}
