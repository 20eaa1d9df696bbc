// Forward-difference formulation of the render inner loop (adds only per pixel) vs the Horner form.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void add_if_le(float& acc, float e, float q, float lim) {
    asm("{\n.reg .pred p;\nsetp.le.f32 p, %2, %3;\n@p add.f32 %0, %0, %1;\n}\n" : "+f"(acc) : "f"(e), "f"(q), "f"(lim));
}
__device__ __forceinline__ void accum_if(float& acc, float w, float e, float q, float lim) {
    asm("{\n.reg .pred p;\nsetp.lt.u32 p, %3, %4;\n@p fma.rn.f32 %0, %1, %2, %0;\n}\n" : "+f"(acc) : "f"(w), "f"(e), "r"(__float_as_uint(q)), "r"(__float_as_uint(lim)));
}
template <int MODE, int PX, int NREC>
__global__ void k(float* out, int iters, const float4* recs) {
    __shared__ float4 s_rec[NREC][2];
    for (int i = threadIdx.x; i < NREC * 2; i += blockDim.x) (&s_rec[0][0])[i] = recs[i];
    __syncthreads();
    const int nsl = (PX == 8) ? 8 : 4;
    const int slice = (PX == 8) ? (threadIdx.x >> 5) : (threadIdx.x >> 6);
    const int q = (PX == 8) ? (threadIdx.x & 31) : (threadIdx.x & 63);
    const float px0 = (PX == 8) ? (float)((q & 1) * 8) : (float)((q & 3) * 4);
    const float py = (PX == 8) ? (float)(q >> 1) : (float)(q >> 2);
    float acc[PX];
    for (int k2 = 0; k2 < PX; ++k2) acc[k2] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int j = slice; j < NREC; j += nsl) {
            const float4 r0 = s_rec[j][0], r1 = s_rec[j][1];
            const float dy = r0.y - py, bdy = r1.y * dy, dx0 = r0.x - px0;
            if (MODE == 0) {
                const float cdy2 = (r1.z * dy) * dy;
#pragma unroll
                for (int k2 = 0; k2 < PX; ++k2) {
                    const float dx = dx0 - (float)k2;
                    const float u = fmaf(r1.x, dx, bdy);
                    const float qq = fmaf(dx, u, cdy2);
                    accum_if(acc[k2], r0.z, ex2(-qq), qq, r0.w);
                }
            } else {
                // q'(k) = A (dx0-k)^2 + bdy (dx0-k) + C dy^2 - log2 w ; r0.z holds log2 w here
                const float cdy2 = fmaf(r1.z * dy, dy, -r0.z);
                const float u0 = fmaf(r1.x, dx0, bdy);
                float qq = fmaf(dx0, u0, cdy2);
                const float a2 = r1.x + r1.x;
                float d = fmaf(-a2, dx0, r1.x - bdy);      // q(k+1) - q(k) at k = 0
                add_if_le(acc[0], ex2(-qq), qq, 16.6096f);
#pragma unroll
                for (int k2 = 1; k2 < PX; ++k2) {
                    qq += d;
                    d += a2;
                    add_if_le(acc[k2], ex2(-qq), qq, 16.6096f);
                }
            }
        }
    }
    float s = 0; for (int k2 = 0; k2 < PX; ++k2) s += acc[k2];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int PX>
void run(const char* name, int sms, int clk, float* out, const float4* d, int bps) {
    const int NREC = 256, iters = 100;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, PX, NREC><<<sms * bps, 256>>>(out, 2, d);
    cudaEventRecord(a);
    k<MODE, PX, NREC><<<sms * bps, 256>>>(out, iters, d);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const int nsl = (PX == 8) ? 8 : 4;
    double mufu = (double)sms * bps * 256 * iters * (NREC / nsl) * PX;
    printf("%-44s %d CTA/SM %.3f ms  %.2f MUFU/SM/clk\n", name, bps, ms, mufu / (ms * 1e-3) / sms / (clk * 1e3));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; float* out; cudaMalloc(&out, sms * 8 * 256 * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int NREC = 256;
    float4 h[NREC * 2];
    for (int j = 0; j < NREC; ++j) { h[2*j] = make_float4(8.f + (j % 7), 8.f - (j % 5), 0.7f, 17.0000019f); h[2*j+1] = make_float4(0.05f, 0.01f, 0.04f, 1.f); }
    float4* d; cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int bps : {4, 6}) {
        run<0, 4>("Horner 4 px", sms, clk, out, d, bps);
        run<1, 4>("forward differences 4 px", sms, clk, out, d, bps);
        run<0, 8>("Horner 8 px", sms, clk, out, d, bps);
        run<1, 8>("forward differences 8 px", sms, clk, out, d, bps);
    }
    return 0;
}
