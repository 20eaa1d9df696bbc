// r2x_voxel.cuh -- launchers for the density-volume (voxelizer) kernels.
#pragma once
#include "r2x_common.cuh"
#include "r2x_binning.cuh"

namespace r2x {

struct VoxelGrid {
    int nx, ny, nz;        // voxels
    float sx, sy, sz;      // physical size
    float cx, cy, cz;      // centre
    int gx, gy, gz;        // 8^3-tile grid
    float dvx, dvy, dvz;   // voxel pitch  (s / n, float32 division)
    float ix, iy, iz;      // 1 / pitch    (float32 reciprocal)
};

struct VoxelGeom {
    float4* rec;              // [4P] (px,py,pz,log2 rho), (a,b,c,d scaled), (e,f scaled, depth, 0 fast | rho exact), (rho,-,-,-)
    uint16_t* cube;           // [6P] x0,y0,z0,x1,y1,z1 tile cube
    uint32_t* tiles_touched;  // [P]
    uint32_t* offsets;        // [P]
};

VoxelGrid make_voxel_grid(int nx, int ny, int nz, float sx, float sy, float sz, float cx, float cy, float cz);

int launch_voxel_preprocess(cudaStream_t st, int P, const float* means, const float* scales, float scale_modifier,
                            const float* rots, const float* opac, const float* cov3D_precomp, const VoxelGrid& vg,
                            int* radii_x, int* radii_y, int* radii_z, const VoxelGeom& geom, const DirectBin* db);
int launch_voxel_render(cudaStream_t st, const VoxelGrid& vg, const VoxelGeom& geom, const uint2* ranges,
                        const uint32_t* point_list, const TilePlan& plan, long long R_launch, float* out_volume);
int launch_voxel_render_bwd(cudaStream_t st, const VoxelGrid& vg, const VoxelGeom& geom, const uint2* ranges,
                            const uint32_t* point_list, const uint32_t* inst_pos, const TilePlan& plan,
                            long long R_launch, const float* dL_dvol, float4* inst_grad);
int launch_voxel_gauss_bwd(cudaStream_t st, int P, const int* radii_x, const int* radii_y, const int* radii_z,
                           const float* scales, float scale_modifier, const float* rots, const float* cov3D_precomp,
                           const VoxelGrid& vg, const VoxelGeom& geom, long long capacity, const uint32_t* inst_pos,
                           const float4* inst_grad, float* dL_dopacity, float* dL_dmean3D, float* dL_dcov3D,
                           float* dL_dscale, float* dL_drot);

}  // namespace r2x
