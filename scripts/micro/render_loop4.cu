// Micro-benchmark: per-pixel inner loop of the forward render kernel, additive vs MULTIPLICATIVE forward differences.
//
//   additive (round 1):   q(k+1) = q(k) + d(k), d(k+1) = d(k) + 2A;  alpha(k) = ex2(-q(k))          -> 1 MUFU / pixel
//   multiplicative:       E(k+1) = E(k) D(k),   D(k+1) = D(k) K;      E(0) = ex2(-q0), D(0) = ex2(-d0), K = ex2(-2A)
//                         -> 2 MUFU per 4-pixel run, FMULs instead of FADDs, test alpha >= 1e-5 directly on E
//   + f32x2: the lane's two 4-pixel runs advance together in packed registers (mul.f32x2 / fma.rn.f32x2)
//   + 16 px per lane: a half warp covers the 16x16 tile, the two half warps take different Gaussians
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o render_loop4 render_loop4.cu && ./render_loop4
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void add_if_le(float& acc, float e, float q, float lim) {
    asm("{\n.reg .pred p;\nsetp.le.f32 p, %2, %3;\n@p add.f32 %0, %0, %1;\n}\n" : "+f"(acc) : "f"(e), "f"(q), "f"(lim));
}
__device__ __forceinline__ void add_if_ge(float& acc, float e) {
    asm("{\n.reg .pred p;\nsetp.ge.f32 p, %1, 0f3727C5AC;\n@p add.f32 %0, %0, %1;\n}\n" : "+f"(acc) : "f"(e));   // 1e-5f
}
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// record: r0 = (x, y, log2 w, -), r1 = (A2, B2, C2, K = 2^(-2 A2))
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, const float4* recs, int nrec) {
    extern __shared__ float4 s_rec[];
    for (int i = threadIdx.x; i < nrec * 2; i += blockDim.x) s_rec[i] = recs[i];
    __syncthreads();
    const int slice = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    if (MODE <= 3) {
        const float px0 = (float)((lane & 1) * 8), py = (float)(lane >> 1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll 2
            for (int j = slice; j < nrec; j += 8) {
                const float4 r0 = s_rec[2 * j], r1 = s_rec[2 * j + 1];
                const float dy = r0.y - py, bdy = r1.y * dy, dx0 = r0.x - px0;
                const float cdy2 = fmaf(r1.z * dy, dy, -r0.z);
                const float a2 = r1.x + r1.x, e0 = r1.x - bdy;
                if (MODE == 0) {            // additive, re-anchored every 4 px (the round-1 kernel)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float dxa = dx0 - (float)(4 * h);
                        float q = fmaf(dxa, fmaf(r1.x, dxa, bdy), cdy2);
                        float d = fmaf(-a2, dxa, e0);
                        add_if_le(acc[4 * h], ex2(-q), q, 16.6096f);
#pragma unroll
                        for (int p = 1; p < 4; ++p) { q += d; d += a2; add_if_le(acc[4 * h + p], ex2(-q), q, 16.6096f); }
                    }
                } else if (MODE == 1) {     // multiplicative, scalar, two runs of 4
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float dxa = dx0 - (float)(4 * h);
                        const float q = fmaf(dxa, fmaf(r1.x, dxa, bdy), cdy2);
                        const float d = fmaf(-a2, dxa, e0);
                        float E = ex2(-q), D = ex2(-d);
                        add_if_ge(acc[4 * h], E);
#pragma unroll
                        for (int p = 1; p < 4; ++p) { E *= D; if (p < 3) D *= r1.w; add_if_ge(acc[4 * h + p], E); }
                    }
                } else if (MODE == 2) {     // multiplicative, the two runs packed (f32x2)
                    const uint64_t DX = pk(dx0, dx0 - 4.0f);
                    const uint64_t U = fma2(pk(r1.x, r1.x), DX, pk(bdy, bdy));
                    const uint64_t Q = fma2(DX, U, pk(cdy2, cdy2));
                    const uint64_t Dd = fma2(pk(-a2, -a2), DX, pk(e0, e0));
                    float q0, q1, d0, d1;
                    upk(Q, q0, q1); upk(Dd, d0, d1);
                    uint64_t E = pk(ex2(-q0), ex2(-q1)), D = pk(ex2(-d0), ex2(-d1));
                    const uint64_t K = pk(r1.w, r1.w);
                    float ea, eb;
                    upk(E, ea, eb); add_if_ge(acc[0], ea); add_if_ge(acc[4], eb);
#pragma unroll
                    for (int p = 1; p < 4; ++p) {
                        E = mul2(E, D);
                        if (p < 3) D = mul2(D, K);
                        upk(E, ea, eb); add_if_ge(acc[p], ea); add_if_ge(acc[4 + p], eb);
                    }
                } else {                    // multiplicative, one run of 8 (speed reference only)
                    const float q = fmaf(dx0, fmaf(r1.x, dx0, bdy), cdy2);
                    const float d = fmaf(-a2, dx0, e0);
                    float E = ex2(-q), D = ex2(-d);
                    add_if_ge(acc[0], E);
#pragma unroll
                    for (int p = 1; p < 8; ++p) { E *= D; if (p < 7) D *= r1.w; add_if_ge(acc[p], E); }
                }
            }
        }
    } else {
        // 16 px per lane: lane & 15 = row, lane >> 4 selects which of two Gaussians; runs of 4, packed in pairs
        const float py = (float)(lane & 15);
        const int sub = lane >> 4;
        for (int it = 0; it < iters; ++it) {
#pragma unroll 2
            for (int j = 2 * slice + sub; j < nrec; j += 16) {
                const float4 r0 = s_rec[2 * j], r1 = s_rec[2 * j + 1];
                const float dy = r0.y - py, bdy = r1.y * dy, dx0 = r0.x;
                const float cdy2 = fmaf(r1.z * dy, dy, -r0.z);
                const float a2 = r1.x + r1.x, e0 = r1.x - bdy;
                const uint64_t K = pk(r1.w, r1.w);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint64_t DX = pk(dx0 - (float)(8 * h), dx0 - (float)(8 * h + 4));
                    const uint64_t U = fma2(pk(r1.x, r1.x), DX, pk(bdy, bdy));
                    const uint64_t Q = fma2(DX, U, pk(cdy2, cdy2));
                    const uint64_t Dd = fma2(pk(-a2, -a2), DX, pk(e0, e0));
                    float q0, q1, d0, d1;
                    upk(Q, q0, q1); upk(Dd, d0, d1);
                    uint64_t E = pk(ex2(-q0), ex2(-q1)), D = pk(ex2(-d0), ex2(-d1));
                    float ea, eb;
                    upk(E, ea, eb); add_if_ge(acc[8 * h], ea); add_if_ge(acc[8 * h + 4], eb);
#pragma unroll
                    for (int p = 1; p < 4; ++p) {
                        E = mul2(E, D);
                        if (p < 3) D = mul2(D, K);
                        upk(E, ea, eb); add_if_ge(acc[8 * h + p], ea); add_if_ge(acc[8 * h + 4 + p], eb);
                    }
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int sms, int clk, float* out, const float4* d, int nrec) {
    for (int bps : {4, 6, 8}) {
        const int iters = 200;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k<MODE><<<sms * bps, 256, nrec * 32>>>(out, 2, d, nrec);
        cudaEventRecord(a);
        k<MODE><<<sms * bps, 256, nrec * 32>>>(out, iters, d, nrec);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        // pixel-Gaussian pairs: every CTA evaluates nrec Gaussians x 256 pixels per iteration
        const double pairs = (double)sms * bps * iters * nrec * 256.0;
        printf("%-46s %d CTA/SM %8.3f ms  %6.2f pairs/clk/SM  (%s)\n", name, bps, ms,
               pairs / (ms * 1e-3) / sms / (clk * 1e3), cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; float* out; cudaMalloc(&out, sms * 8 * 256 * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int NREC = 256;
    float4 h[NREC * 2];
    for (int j = 0; j < NREC; ++j) {
        const float A2 = 0.05f + 0.001f * (j % 11);
        h[2 * j] = make_float4(8.f + (j % 7), 8.f - (j % 5), -0.7f, 0.f);
        h[2 * j + 1] = make_float4(A2, 0.01f, 0.04f, exp2f(-2.f * A2));
    }
    float4* d; cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("device %s, %d SMs, %d MHz nominal; MUFU floor = 16 pairs/clk/SM for the additive form\n", p.name, sms, clk / 1000);
    run<0>("additive fwd-diff, 2 runs of 4 (round 1)", sms, clk, out, d, NREC);
    run<1>("multiplicative, 2 runs of 4, scalar", sms, clk, out, d, NREC);
    run<2>("multiplicative, 2 runs of 4, f32x2", sms, clk, out, d, NREC);
    run<3>("multiplicative, 1 run of 8, scalar", sms, clk, out, d, NREC);
    run<4>("multiplicative, 16 px/lane, f32x2", sms, clk, out, d, NREC);
    return 0;
}
