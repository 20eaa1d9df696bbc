// ref_shim.cu -- C-ABI shim around the UNMODIFIED reference CUDA rasterizer / voxelizer.
//
// TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product library).
// oracle/build_ref.sh compiles the reference's own six .cu files from where they lie under
// /root/reference/.../xray-gaussian-rasterization-voxelization/ together with this file into
// oracle/_ref/libr2ref.so.  The shim only (a) provides the three growable scratch buffers the
// reference asks for through std::function callbacks (SUB/utility.h:7-13 does the same with torch
// tensors), (b) forwards the raw device pointers to CudaRasterizer::Rasterizer::{forward,backward}
// (RAS/rasterizer.h:37-85) and CudaVoxelizer::Voxelizer::{forward,backward} (VOX/voxelizer.h:28-72),
// and (c) copies the reference's internal state arrays out (layout from RAS/rasterizer_impl.h:29-63,
// VOX/voxelizer_impl.h:30-65) so tests can compare stage by stage.  All pointers are device pointers.
#include <cstdint>
#include <cstdio>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"
#include "cuda_voxelizer/voxelizer.h"
#include "cuda_voxelizer/voxelizer_impl.h"

namespace {
struct Grow {
    char* p = nullptr;
    size_t cap = 0;
    char* get(size_t n) {
        if (n > cap) {
            if (p) cudaFree(p);
            size_t want = n + n / 4 + 4096;
            if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; cap = 0; return nullptr; }
            cap = want;
        }
        return p;
    }
};
Grow g_ras[3], g_vox[3];
inline void d2d(void* dst, const void* src, size_t n) {
    if (dst && n) cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice);
}
}  // namespace

extern "C" {

int ref_raster_forward(int P, int W, int H, const float* means3D, const float* opacities, const float* scales,
                       float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* view, const float* proj, const float* campos, float tan_fovx,
                       float tan_fovy, int mode, float* out_color, int* radii) {
    std::function<char*(size_t)> f0 = [](size_t n) { return g_ras[0].get(n); };
    std::function<char*(size_t)> f1 = [](size_t n) { return g_ras[1].get(n); };
    std::function<char*(size_t)> f2 = [](size_t n) { return g_ras[2].get(n); };
    return CudaRasterizer::Rasterizer::forward(f0, f1, f2, P, W, H, means3D, opacities, scales, scale_modifier,
                                               rotations, cov3D_precomp, view, proj, campos, tan_fovx, tan_fovy,
                                               false, mode, out_color, radii, false);
}

void ref_raster_export(int P, int W, int H, int R, float* depths, float* means2D, float* cov3D,
                       float* conic_opacity, float* mus, uint32_t* tiles_touched, uint32_t* point_offsets,
                       uint64_t* keys_unsorted, uint32_t* vals_unsorted, uint64_t* keys_sorted,
                       uint32_t* point_list, uint32_t* ranges, uint32_t* n_contrib) {
    char* c = g_ras[0].p;
    auto geom = CudaRasterizer::GeometryState::fromChunk(c, P);
    d2d(depths, geom.depths, sizeof(float) * P);
    d2d(means2D, geom.means2D, sizeof(float) * 2 * P);
    d2d(cov3D, geom.cov3D, sizeof(float) * 6 * P);
    d2d(conic_opacity, geom.conic_opacity, sizeof(float) * 4 * P);
    d2d(mus, geom.mus, sizeof(float) * P);
    d2d(tiles_touched, geom.tiles_touched, sizeof(uint32_t) * P);
    d2d(point_offsets, geom.point_offsets, sizeof(uint32_t) * P);
    c = g_ras[1].p;
    auto bin = CudaRasterizer::BinningState::fromChunk(c, R);
    d2d(keys_unsorted, bin.point_list_keys_unsorted, sizeof(uint64_t) * R);
    d2d(vals_unsorted, bin.point_list_unsorted, sizeof(uint32_t) * R);
    d2d(keys_sorted, bin.point_list_keys, sizeof(uint64_t) * R);
    d2d(point_list, bin.point_list, sizeof(uint32_t) * R);
    c = g_ras[2].p;
    auto img = CudaRasterizer::ImageState::fromChunk(c, (size_t)W * H);
    int tiles = ((W + 15) / 16) * ((H + 15) / 16);
    d2d(ranges, img.ranges, sizeof(uint32_t) * 2 * tiles);
    d2d(n_contrib, img.n_contrib, sizeof(uint32_t) * (size_t)W * H);
    cudaDeviceSynchronize();
}

void ref_raster_backward(int P, int R, int W, int H, const float* means3D, const float* scales,
                         float scale_modifier, const float* rotations, const float* cov3D_precomp,
                         const float* view, const float* proj, const float* campos, float tan_fovx,
                         float tan_fovy, const int* radii, const float* dL_dpix, float* dL_dmean2D,
                         float* dL_dconic, float* dL_dopacity, float* dL_dmu, float* dL_dmean3D,
                         float* dL_dcov3D, float* dL_dscale, float* dL_drot, int mode) {
    CudaRasterizer::Rasterizer::backward(P, R, W, H, means3D, scales, scale_modifier, rotations, cov3D_precomp,
                                         view, proj, campos, tan_fovx, tan_fovy, radii, g_ras[0].p, g_ras[1].p,
                                         g_ras[2].p, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dmu,
                                         dL_dmean3D, dL_dcov3D, dL_dscale, dL_drot, mode, false);
}

void ref_mark_visible(int P, float* means3D, float* view, float* proj, bool* present) {
    CudaRasterizer::Rasterizer::markVisible(P, means3D, view, proj, present);
}

int ref_voxel_forward(int P, int nx, int ny, int nz, float sx, float sy, float sz, float cx, float cy, float cz,
                      const float* means3D, const float* opacities, const float* scales, float scale_modifier,
                      const float* rotations, const float* cov3D_precomp, float* out_volume, int* radii_x,
                      int* radii_y, int* radii_z) {
    std::function<char*(size_t)> f0 = [](size_t n) { return g_vox[0].get(n); };
    std::function<char*(size_t)> f1 = [](size_t n) { return g_vox[1].get(n); };
    std::function<char*(size_t)> f2 = [](size_t n) { return g_vox[2].get(n); };
    return CudaVoxelizer::Voxelizer::forward(f0, f1, f2, P, nx, ny, nz, sx, sy, sz, cx, cy, cz, means3D, opacities,
                                             scales, scale_modifier, rotations, cov3D_precomp, false, out_volume,
                                             radii_x, radii_y, radii_z, false);
}

void ref_voxel_export(int P, int nx, int ny, int nz, int R, float* depths, float* means3D_norm, float* cov3D,
                      float* conic_opacity, uint32_t* tiles_touched, uint32_t* point_offsets,
                      uint64_t* keys_unsorted, uint32_t* vals_unsorted, uint64_t* keys_sorted,
                      uint32_t* point_list, uint32_t* ranges, uint32_t* n_contrib) {
    char* c = g_vox[0].p;
    auto geom = CudaVoxelizer::GeometryState::fromChunk(c, P);
    d2d(depths, geom.depths, sizeof(float) * P);
    d2d(means3D_norm, geom.means3D_norm, sizeof(float) * 3 * P);
    d2d(cov3D, geom.cov3D, sizeof(float) * 6 * P);
    d2d(conic_opacity, geom.conic_opacity, sizeof(float) * 7 * P);
    d2d(tiles_touched, geom.tiles_touched, sizeof(uint32_t) * P);
    d2d(point_offsets, geom.point_offsets, sizeof(uint32_t) * P);
    c = g_vox[1].p;
    auto bin = CudaVoxelizer::BinningState::fromChunk(c, R);
    d2d(keys_unsorted, bin.point_list_keys_unsorted, sizeof(uint64_t) * R);
    d2d(vals_unsorted, bin.point_list_unsorted, sizeof(uint32_t) * R);
    d2d(keys_sorted, bin.point_list_keys, sizeof(uint64_t) * R);
    d2d(point_list, bin.point_list, sizeof(uint32_t) * R);
    c = g_vox[2].p;
    size_t N = (size_t)nx * ny * nz;
    auto img = CudaVoxelizer::ImageState::fromChunk(c, N);
    size_t tiles = (size_t)((nx + 7) / 8) * ((ny + 7) / 8) * ((nz + 7) / 8);
    d2d(ranges, img.ranges, sizeof(uint32_t) * 2 * tiles);
    d2d(n_contrib, img.n_contrib, sizeof(uint32_t) * N);
    cudaDeviceSynchronize();
}

void ref_voxel_backward(int P, int R, int nx, int ny, int nz, float sx, float sy, float sz, float cx, float cy,
                        float cz, const float* means3D, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp, const int* radii_x,
                        const int* radii_y, const int* radii_z, const float* dL_dvol, float* dL_dmean3D_norm,
                        float* dL_dconic3D, float* dL_dopacity, float* dL_dmean3D, float* dL_dcov3D,
                        float* dL_dscale, float* dL_drot) {
    CudaVoxelizer::Voxelizer::backward(P, R, nx, ny, nz, sx, sy, sz, cx, cy, cz, means3D, scales, scale_modifier,
                                       rotations, cov3D_precomp, radii_x, radii_y, radii_z, g_vox[0].p, g_vox[1].p,
                                       g_vox[2].p, dL_dvol, dL_dmean3D_norm, dL_dconic3D, dL_dopacity, dL_dmean3D,
                                       dL_dcov3D, dL_dscale, dL_drot, false);
}

int ref_last_cuda_error(void) { return (int)cudaGetLastError(); }

}  // extern "C"
