// Issue-rate microbenchmark for the packed FP32 instructions of sm_100a (FADD2 / FFMA2) against their scalar forms, alone
// and interleaved with integer ALU work.  Question: does one FFMA2 (64 lane-operations) cost one issue slot or two?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_issue profiles/micro/f32x2_issue.cu ; run: ./f32x2_issue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fadd1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned d; asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, int iters, float s)
{
    const int N = 8;
    float a[2 * N]; u64 p[N]; unsigned q[N];
    for (int i = 0; i < 2 * N; i++) a[i] = s * (threadIdx.x + i);
    for (int i = 0; i < N; i++) { p[i] = ((u64)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]); q[i] = threadIdx.x * 7 + i; }
    const u64 cb = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s + 1.f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (MODE == 0) { a[i] = ffma1(a[i], s, a[i + N]); a[i + N] = ffma1(a[i + N], s, a[i]); }            // 2 FFMA
            if (MODE == 1) p[i] = ffma2(p[i], cb, p[(i + 1) % N]);                                                 // 1 FFMA2 (same lane-ops as mode 0)
            if (MODE == 2) { a[i] = ffma1(a[i], s, a[i + N]); a[i + N] = ffma1(a[i + N], s, a[i]); q[i] = lop(q[i], q[(i + 1) % N]); q[i] = lop(q[i], it); }   // 2 FFMA + 2 LOP
            if (MODE == 3) { p[i] = ffma2(p[i], cb, p[(i + 1) % N]); q[i] = lop(q[i], q[(i + 1) % N]); q[i] = lop(q[i], it); }   // 1 FFMA2 + 2 LOP
            if (MODE == 4) { a[i] = fadd1(a[i], a[i + N]); a[i + N] = fadd1(a[i + N], s); }                       // 2 FADD
            if (MODE == 5) p[i] = fadd2(p[i], p[(i + 1) % N]);                                                     // 1 FADD2
            if (MODE == 6) { q[i] = lop(q[i], q[(i + 1) % N]); q[i] = lop(q[i], it); }                             // 2 LOP
            if (MODE == 7) { p[i] = ffma2(p[i], cb, p[(i + 1) % N]); q[i] = lop(q[i], q[(i + 1) % N]); }          // 1 FFMA2 + 1 LOP
        }
    }
    float r = 0.f;
    for (int i = 0; i < 2 * N; i++) r += a[i];
    for (int i = 0; i < N; i++) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + (float)q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> static void run(const char *name, double lane_ops_per_iter, double instr_per_iter)
{
    float *out; cudaMalloc(&out, 148 * 4 * 512 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<148 * 4, 512>>>(out, iters, 1.0001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double warps = 148.0 * 4 * 16, winst = warps * iters * 8 * instr_per_iter;
    const double cyc = best * 1e-3 * 1.965e9;       // assumes 1965 MHz
    printf("{\"mode\": \"%s\", \"ms\": %.3f, \"warp_instr_per_clk_per_smsp\": %.3f, \"fp32_lane_ops_per_clk_per_sm\": %.1f}\n", name, best,
           winst / cyc / (148 * 4), warps * iters * 8 * lane_ops_per_iter * 32 / cyc / 148);
    cudaFree(out);
}
int main()
{
    run<0>("2 FFMA", 2, 2); run<1>("1 FFMA2", 2, 1); run<2>("2 FFMA + 2 LOP", 2, 4); run<3>("1 FFMA2 + 2 LOP", 2, 3);
    run<4>("2 FADD", 2, 2); run<5>("1 FADD2", 2, 1); run<6>("2 LOP", 0, 2); run<7>("1 FFMA2 + 1 LOP", 2, 2);
    return 0;
}
