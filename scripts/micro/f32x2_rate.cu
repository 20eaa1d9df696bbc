// Micro-benchmark: packed FP32 (add.rn.f32x2 / fma.rn.f32x2, sm_100) vs scalar FADD / FFMA issue rate, and the
// render inner loop written with packed forward differences.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; b[i] = 1e-6f * (i + 1); }
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fadd(a[i], b[i]);
        }
    } else if (MODE == 1) {
        uint64_t A[4], B[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { A[i] = pk(a[2 * i], a[2 * i + 1]); B[i] = pk(b[2 * i], b[2 * i + 1]); }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) A[i] = add2(A[i], B[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) upk(A[i], a[2 * i], a[2 * i + 1]);
    } else if (MODE == 2) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ffma(a[i], b[i], b[(i + 1) & 7]);
        }
    } else if (MODE == 3) {
        uint64_t A[4], B[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { A[i] = pk(a[2 * i], a[2 * i + 1]); B[i] = pk(b[2 * i], b[2 * i + 1]); }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) A[i] = fma2(A[i], B[i], B[(i + 1) & 3]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) upk(A[i], a[2 * i], a[2 * i + 1]);
    } else if (MODE == 4) {   // current render loop, 8 px per "Gaussian": scalar forward differences
        for (int it = 0; it < iters; ++it) {
            float q = a[0], d = b[0];
            const float a2 = b[1];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const float e = ex2(-q);
                if (q <= 16.6f) acc[p] += e;
                q = fadd(q, d);
                d = fadd(d, a2);
            }
            a[0] += 1e-3f;
        }
    } else if (MODE == 5) {   // packed forward differences: pixel pairs
        for (int it = 0; it < iters; ++it) {
            uint64_t Q = pk(a[0], a[0] + b[0]), D = pk(b[0] + b[0], b[0] + b[0] + b[1]);
            const uint64_t DD = pk(b[1], b[1]);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float q0, q1;
                upk(Q, q0, q1);
                const float e0 = ex2(-q0), e1 = ex2(-q1);
                if (q0 <= 16.6f) acc[2 * p] += e0;
                if (q1 <= 16.6f) acc[2 * p + 1] += e1;
                Q = add2(Q, D);
                D = add2(D, DD);
            }
            a[0] += 1e-3f;
        }
    } else if (MODE == 6) {   // packed, accumulate with packed add of selected values
        uint64_t ACC[4] = {0, 0, 0, 0};
        for (int it = 0; it < iters; ++it) {
            uint64_t Q = pk(a[0], a[0] + b[0]), D = pk(b[0] + b[0], b[0] + b[0] + b[1]);
            const uint64_t DD = pk(b[1], b[1]);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float q0, q1;
                upk(Q, q0, q1);
                float e0 = ex2(-q0), e1 = ex2(-q1);
                e0 = (q0 <= 16.6f) ? e0 : 0.f;
                e1 = (q1 <= 16.6f) ? e1 : 0.f;
                ACC[p] = add2(ACC[p], pk(e0, e1));
                Q = add2(Q, D);
                D = add2(D, DD);
            }
            a[0] += 1e-3f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) upk(ACC[i], acc[2 * i], acc[2 * i + 1]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* out, int sms, int clk) {
    for (int bps : {4, 6, 8}) {
        const int iters = 20000;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k<MODE><<<sms * bps, 256>>>(out, 100, 0.1f);
        cudaEventRecord(a);
        k<MODE><<<sms * bps, 256>>>(out, iters, 0.1f);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double ops = (double)sms * bps * 256 * iters * 8;   // 8 float lane-results (or pixels) per iteration
        printf("%-34s %d CTAs/SM: %.3f ms  %.1f lane-results/clk/SM (@%d MHz nominal)\n", name, bps, ms,
               ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; float* out; cudaMalloc(&out, sms * 8 * 256 * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    run<0>("FADD scalar", out, sms, clk);
    run<1>("add.rn.f32x2", out, sms, clk);
    run<2>("FFMA scalar 3-reg", out, sms, clk);
    run<3>("fma.rn.f32x2", out, sms, clk);
    run<4>("render px loop scalar fwd-diff", out, sms, clk);
    run<5>("render px loop packed fwd-diff", out, sms, clk);
    run<6>("render px loop packed + packed acc", out, sms, clk);
    return 0;
}
