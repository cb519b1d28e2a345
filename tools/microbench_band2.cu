// Microbenchmark, round 2: what bounds the min-plus band of the tensor-core DTW scan, and which reformulations move it.
//   A   fp32 band, one segment per thread            FMNMX3 + FADD per cell                      (round 1: 45.3 cells/clk/SM)
//   B   fp32, TWO independent bands per thread with their columns interleaved in registers, cost adds packed as add.f32x2:
//       2 FMNMX3 + 1 FADD2 per cell pair
//   C   half2: two independent bands in the two halves of a register: VHMNMX (3-input min) + HADD2 per cell pair
//   T*  raw issue rates of the instructions involved (independent chains)
// Reports cells / clk / SM at the measured SM clock (clock64 deltas). Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_band2 microbench_band2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long x, y, z;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
    asm("add.f32x2 %0, %1, %2;" : "=l"(z) : "l"(x), "l"(y));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(z));
    return r;
}

// ---- A: fp32, one band --------------------------------------------------------------------------------------------------
template <int R, int NC, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_a(const float* __restrict__ in, float* __restrict__ out, int steps, long long* clk) {
    const float INF = __int_as_float(0x7f800000);
    float tm[R][NC], d[NC];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < NC; j++) tm[r][j] = in[(r * NC + j) * 32 + (threadIdx.x & 31)];
#pragma unroll
    for (int j = 0; j < NC; j++) d[j] = INF;
    float dinit = 0.f;
    const long long t0 = clock64();
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
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NC; j++) s += d[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

// ---- B: fp32, two bands, interleaved columns, FADD2 -------------------------------------------------------------------
template <int R, int NC, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_b(const float* __restrict__ in, float* __restrict__ out, int steps, long long* clk) {
    const float INF = __int_as_float(0x7f800000);
    float2 tm[R][NC], d[NC];  // .x = band A, .y = band B
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < NC; j++) tm[r][j] = make_float2(in[(r * NC + j) * 32 + (threadIdx.x & 31)], in[(r * NC + j) * 32 + ((threadIdx.x + 1) & 31)]);
#pragma unroll
    for (int j = 0; j < NC; j++) d[j] = make_float2(INF, INF);
    float dinit = 0.f;
    const long long t0 = clock64();
    for (int st = 0; st < steps; st++) {
        float2 left[R], diag0 = make_float2(dinit, dinit);
#pragma unroll
        for (int r = 0; r < R; r++) left[r] = make_float2(INF, INF);
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const float2 up0 = d[j];
            float2 c[R];
            c[0] = add2(tm[0][j], make_float2(min3(left[0].x, up0.x, diag0.x), min3(left[0].y, up0.y, diag0.y)));
#pragma unroll
            for (int r = 1; r < R; r++)
                c[r] = add2(tm[r][j], make_float2(min3(left[r].x, c[r - 1].x, left[r - 1].x), min3(left[r].y, c[r - 1].y, left[r - 1].y)));
            diag0 = up0;
#pragma unroll
            for (int r = 0; r < R; r++) left[r] = c[r];
            d[j] = c[R - 1];
        }
        dinit = INF;
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NC; j++) s += d[j].x + d[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

// ---- C: half2, two bands in the halves --------------------------------------------------------------------------------
template <int R, int NC, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_c(const float* __restrict__ in, float* __restrict__ out, int steps, long long* clk) {
    const __half2 INF = __float2half2_rn(60000.f);
    __half2 tm[R][NC], d[NC];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < NC; j++) tm[r][j] = __float2half2_rn(in[(r * NC + j) * 32 + (threadIdx.x & 31)]);
#pragma unroll
    for (int j = 0; j < NC; j++) d[j] = INF;
    __half2 dinit = __float2half2_rn(0.f);
    const long long t0 = clock64();
    for (int st = 0; st < steps; st++) {
        __half2 left[R], diag0 = dinit;
#pragma unroll
        for (int r = 0; r < R; r++) left[r] = INF;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const __half2 up0 = d[j];
            __half2 c[R];
            c[0] = __hadd2(tm[0][j], __hmin2(__hmin2(left[0], up0), diag0));
#pragma unroll
            for (int r = 1; r < R; r++) c[r] = __hadd2(tm[r][j], __hmin2(__hmin2(left[r], c[r - 1]), left[r - 1]));
            diag0 = up0;
#pragma unroll
            for (int r = 0; r < R; r++) left[r] = c[r];
            d[j] = c[R - 1];
        }
        dinit = INF;
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NC; j++) s += __low2float(d[j]) + __high2float(d[j]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

// ---- T: raw issue rates: 16 independent chains per thread --------------------------------------------------------------
template <int OP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_t(const float* __restrict__ in, float* __restrict__ out, int steps, long long* clk) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = in[j * 32 + (threadIdx.x & 31)];
    const float a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    const long long t0 = clock64();
    for (int st = 0; st < steps; st++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (OP == 0) v[j] = min3(v[j], a, b);                     // FMNMX3
            if (OP == 1) v[j] = v[j] + a;                              // FADD
            if (OP == 2) v[j] = fminf(v[j], a);                        // FMNMX
            if (OP == 3) {                                             // VHMNMX
                __half2 h = *reinterpret_cast<__half2*>(&v[j]);
                h = __hmin2(__hmin2(h, *reinterpret_cast<const __half2*>(&a)), *reinterpret_cast<const __half2*>(&b));
                v[j] = *reinterpret_cast<float*>(&h);
            }
            if (OP == 4) {                                             // HADD2
                __half2 h = *reinterpret_cast<__half2*>(&v[j]);
                h = __hadd2(h, *reinterpret_cast<const __half2*>(&a));
                v[j] = *reinterpret_cast<float*>(&h);
            }
            if (OP == 5 && (j & 1) == 0) {                             // FADD2 (8 per 16 values)
                float2 r = add2(make_float2(v[j], v[j + 1]), make_float2(a, b));
                v[j] = r.x, v[j + 1] = r.y;
            }
            if (OP == 6) v[j] = (j & 1) ? v[j] + a : min3(v[j], a, b);  // FMNMX3 and FADD alternating, independent
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j++) s += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <typename K>
static int timeit(K kern, const char* name, int warps, int sms, float* in, float* out, int steps, double cells_per_thread, long long* d_clk) {
    kern<<<sms, warps * 32>>>(in, out, steps, d_clk);
    CK(cudaDeviceSynchronize());
    kern<<<sms, warps * 32>>>(in, out, steps, d_clk);
    CK(cudaDeviceSynchronize());
    long long clk = 0;
    CK(cudaMemcpy(&clk, d_clk, sizeof(clk), cudaMemcpyDeviceToHost));
    printf("%-44s warps/SM=%2d: %9lld clk  %.2f /clk/SM\n", name, warps, clk, cells_per_thread * warps * 32 / (double)clk);
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float *in, *out;
    long long* d_clk;
    CK(cudaMalloc(&in, 16 * 64 * 32 * sizeof(float)));
    CK(cudaMemset(in, 0, 16 * 64 * 32 * sizeof(float)));
    CK(cudaMalloc(&out, (size_t)sms * 1024 * sizeof(float)));
    CK(cudaMalloc(&d_clk, 8));
    const int S = 2048;
    for (int warps : {8, 16}) {
#define RUN_A(R, NC, MT) if (timeit(k_a<R, NC, MT>, "A fp32 1 band R=" #R " NC=" #NC, warps, sms, in, out, S / R, (double)(S / R) * R * NC, d_clk)) return 1;
#define RUN_B(R, NC, MT) if (timeit(k_b<R, NC, MT>, "B fp32 2 bands FADD2 R=" #R " NC=" #NC "x2", warps, sms, in, out, S / R, (double)(S / R) * R * NC * 2, d_clk)) return 1;
#define RUN_C(R, NC, MT) if (timeit(k_c<R, NC, MT>, "C half2 2 bands R=" #R " NC=" #NC "x2", warps, sms, in, out, S / R, (double)(S / R) * R * NC * 2, d_clk)) return 1;
        if (warps == 8) {
            RUN_A(2, 32, 256) RUN_A(4, 32, 256) RUN_B(2, 16, 256) RUN_B(2, 32, 256) RUN_B(4, 16, 256) RUN_C(2, 32, 256) RUN_C(4, 32, 256) RUN_C(4, 16, 256)
        } else {
            RUN_A(2, 32, 512) RUN_A(4, 16, 512) RUN_B(2, 16, 512) RUN_B(4, 8, 512) RUN_C(2, 32, 512) RUN_C(4, 16, 512) RUN_C(2, 16, 512)
        }
    }
    for (int warps : {4, 8, 16}) {
        if (timeit(k_t<0, 512>, "T FMNMX3 (per instr)", warps, sms, in, out, S, (double)S * 16, d_clk)) return 1;
        if (timeit(k_t<1, 512>, "T FADD", warps, sms, in, out, S, (double)S * 16, d_clk)) return 1;
        if (timeit(k_t<2, 512>, "T FMNMX", warps, sms, in, out, S, (double)S * 16, d_clk)) return 1;
        if (timeit(k_t<3, 512>, "T VHMNMX (per instr = 2 values)", warps, sms, in, out, S, (double)S * 16, d_clk)) return 1;
        if (timeit(k_t<4, 512>, "T HADD2 (per instr = 2 values)", warps, sms, in, out, S, (double)S * 16, d_clk)) return 1;
        if (timeit(k_t<5, 512>, "T FADD2 (per instr = 2 values)", warps, sms, in, out, S, (double)S * 8, d_clk)) return 1;
        if (timeit(k_t<6, 512>, "T FMNMX3 + FADD alternating (per instr)", warps, sms, in, out, S, (double)S * 16, d_clk)) return 1;
    }
    printf("done\n");
    return 0;
}
