// Microbenchmark behind the tensor-core DTW scan's warp / band shape: the register-resident min-plus
// band alone (FMNMX3 + FADD per cell, costs already in registers), for R rows per band (ILP R) and W warps per SM.
// Reports cells / clk / SM; the ALU pipe (FMNMX3, 16 lanes / clk / SMSP) caps it at 64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_band microbench_band.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int R, int NC, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_band(const float* __restrict__ in, float* __restrict__ out, int steps) {
    const float INF = __int_as_float(0x7f800000);
    float tm[R][NC], d[NC];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < NC; j++) tm[r][j] = in[(r * NC + j) * 32 + (threadIdx.x & 31)];
#pragma unroll
    for (int j = 0; j < NC; j++) d[j] = INF;
    float dinit = 0.f;
    for (int st = 0; st < steps; st++) {
        float left[R], diag0 = dinit;
#pragma unroll
        for (int r = 0; r < R; r++) left[r] = INF;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const float up0 = d[j];
            float c[R];
            c[0] = tm[0][j] + min3(left[0], up0, diag0);
#pragma unroll
            for (int r = 1; r < R; r++) c[r] = tm[r][j] + min3(left[r], c[r - 1], left[r - 1]);
            diag0 = up0;
#pragma unroll
            for (int r = 0; r < R; r++) left[r] = c[r];
            d[j] = c[R - 1];
        }
        dinit = INF;
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NC; j++) s += d[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int R, int NC, int MAXT>
static int run1(int warps, int sms, float* in, float* out) {
    const int steps = 4096 / R;
    auto kern = k_band<R, NC, MAXT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    kern<<<sms, warps * 32, 200 * 1024>>>(in, out, steps);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int i = 0; i < reps; i++) kern<<<sms, warps * 32, 200 * 1024>>>(in, out, steps);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double cells = (double)sms * warps * 32 * steps * R * NC;
    printf("R=%d NC=%2d warps/SM=%2d: %.3f ms  %.2f cells/clk/SM\n", R, NC, warps, ms, cells / (ms * 1e-3) / sms / 1.965e9);
    return 0;
}

template <int R, int NC>
static int run(int warps, int sms, float* in, float* out) {
    if (warps <= 8) return run1<R, NC, 256>(warps, sms, in, out);
    if (warps <= 12) return run1<R, NC, 384>(warps, sms, in, out);
    return run1<R, NC, 512>(warps, sms, in, out);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float *in, *out;
    CK(cudaMalloc(&in, 8 * 32 * 32 * sizeof(float)));
    CK(cudaMemset(in, 0, 8 * 32 * 32 * sizeof(float)));
    CK(cudaMalloc(&out, (size_t)sms * 1024 * sizeof(float)));
    for (int warps : {4, 8, 12, 16}) {
        if (run<1, 32>(warps, sms, in, out)) return 1;
        if (run<2, 32>(warps, sms, in, out)) return 1;
        if (run<2, 16>(warps, sms, in, out)) return 1;
        if (warps <= 12 && run<4, 32>(warps, sms, in, out)) return 1;
        if (run<4, 16>(warps, sms, in, out)) return 1;
        if (warps <= 8 && run<6, 32>(warps, sms, in, out)) return 1;
        if (run<8, 16>(warps, sms, in, out)) return 1;
    }
    printf("done\n");
    return 0;
}
