// Micro-benchmark of the render inner loop in isolation: records broadcast from shared memory (LDS.128 x2),
// 4 pixels per thread, the exact arithmetic of raster_render_kernel.  Reports MUFU lane-ops per SM per clock.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void accum_if(float& acc, float w, float e, float q, float lim) {
    asm("{\n.reg .pred p;\nsetp.lt.u32 p, %3, %4;\n@p fma.rn.f32 %0, %1, %2, %0;\n}\n" : "+f"(acc) : "f"(w), "f"(e), "r"(__float_as_uint(q)), "r"(__float_as_uint(lim)));
}
template <int PX, int NREC>
__global__ void k(float* out, int iters, const float4* recs) {
    __shared__ float4 s_rec[NREC][2];
    for (int i = threadIdx.x; i < NREC * 2; i += blockDim.x) (&s_rec[0][0])[i] = recs[i];
    __syncthreads();
    const int slice = threadIdx.x >> 6, q = threadIdx.x & 63;
    const float px0 = (float)((q & 3) * 4), py = (float)(q >> 2);
    float acc[PX];
    for (int k2 = 0; k2 < PX; ++k2) acc[k2] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int j = slice; j < NREC; j += 4) {
            const float4 r0 = s_rec[j][0], r1 = s_rec[j][1];
            const float dy = r0.y - py, bdy = r1.y * dy, cdy2 = (r1.z * dy) * dy, dx0 = r0.x - px0;
#pragma unroll
            for (int k2 = 0; k2 < PX; ++k2) {
                const float dx = dx0 - (float)k2;
                const float u = fmaf(r1.x, dx, bdy);
                const float qq = fmaf(dx, u, cdy2);
                accum_if(acc[k2], r0.z, ex2(-qq), qq, r0.w);
            }
        }
    }
    float s = 0; for (int k2 = 0; k2 < PX; ++k2) s += acc[k2];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; float* out; cudaMalloc(&out, sms * 8 * 256 * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int NREC = 256;
    float4 h[NREC * 2];
    for (int j = 0; j < NREC; ++j) { h[2*j] = make_float4(8.f + (j % 7), 8.f - (j % 5), 0.7f, 17.0000019f); h[2*j+1] = make_float4(0.05f, 0.01f, 0.04f, 1.f); }
    float4* d; cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int bps : {2, 4, 6, 8}) {
        const int iters = 200;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k<4, NREC><<<sms * bps, 256>>>(out, 2, d);
        cudaEventRecord(a);
        k<4, NREC><<<sms * bps, 256>>>(out, iters, d);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double mufu = (double)sms * bps * 256 * iters * (NREC / 4) * 4;
        printf("4px/thread, %d CTAs/SM: %.3f ms, %.2f MUFU lane-ops per SM per clk (nominal %d MHz)\n", bps, ms, mufu / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    }
    return 0;
}
