/*
 * r2x.h -- C ABI of libr2xray.so, the B200-native X-ray Gaussian rasterizer + voxelizer.
 *
 * Drop-in boundary for the reference's native entry points (plain pointers and sizes, no torch types):
 *
 *   r2x_raster_forward / _async  <->  CudaRasterizer::Rasterizer::forward   (RAS/rasterizer.h:37-58,
 *                                     bound by RasterizeGaussiansCUDA, SUB/rasterize_points.cu:28-97)
 *   r2x_raster_backward          <->  CudaRasterizer::Rasterizer::backward  (RAS/rasterizer.h:60-85,
 *                                     RasterizeGaussiansBackwardCUDA, SUB/rasterize_points.cu:99-164)
 *   r2x_mark_visible             <->  CudaRasterizer::Rasterizer::markVisible (RAS/rasterizer.h:30-35,
 *                                     markVisible, SUB/rasterize_points.cu:166-186)
 *   r2x_voxel_forward / _async   <->  CudaVoxelizer::Voxelizer::forward     (VOX/voxelizer.h:28-49,
 *                                     VoxelizeGaussiansCUDA, SUB/voxelize_points.cu:29-98)
 *   r2x_voxel_backward           <->  CudaVoxelizer::Voxelizer::backward    (VOX/voxelizer.h:51-72,
 *                                     VoxelizeGaussiansBackwardCUDA, SUB/voxelize_points.cu:102-167)
 *   r2x_*_export                 --   stage outputs for parity tests (the reference keeps them inside
 *                                     its opaque geom/binning/img byte buffers, RAS/rasterizer_impl.h:29-63)
 *
 * (RAS/VOX/SUB = r2_gaussian/submodules/xray-gaussian-rasterization-voxelization/{cuda_rasterizer,
 *  cuda_voxelizer,.} in the reference.)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked "host";
 *     float32, contiguous, same layouts as the reference: means3D[P,3], scales[P,3], rotations[P,4]
 *     (r,x,y,z, consumed un-normalised), opacities[P], cov3D_precomp[P,6] or NULL, viewmatrix /
 *     projmatrix 16 floats column-major-flat (the reference passes them transposed), out_color[1,H,W],
 *     out_volume[nx,ny,nz] (index x*ny*nz + y*nz + z), radii int32[P].
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is enqueued
 *     on it; the *_async entry points never synchronise with the host.
 *   - the three state buffers play the role of the reference's geomBuffer / binningBuffer / imgBuffer:
 *     opaque to the caller, sized by the r2x_*_bytes functions, written by forward, read by backward.
 *   - every function returns 0 on success; otherwise a non-zero code with r2x_last_error() describing
 *     it (invalid argument, CUDA error, capacity overflow).  No CPU fallback exists.
 */
#ifndef R2X_H_INCLUDED
#define R2X_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R2X_OK 0
#define R2X_ERR_INVALID 1
#define R2X_ERR_CUDA 2
#define R2X_ERR_OVERFLOW 3

/* Allocator callback used by the synchronous forward calls for the binning buffer, whose size depends
 * on the instance count R that is only known mid-forward (the reference does the same through
 * std::function<char*(size_t)>, SUB/utility.h:7-13).  Must return a device pointer to >= nbytes. */
typedef void* (*r2x_alloc_fn)(size_t nbytes, void* user);

const char* r2x_last_error(void);
int r2x_version(void);

/* ---- buffer sizes -------------------------------------------------------------------------- */
size_t r2x_raster_geom_bytes(int P);
size_t r2x_raster_image_bytes(int P, int W, int H);
size_t r2x_voxel_geom_bytes(int P);
size_t r2x_voxel_image_bytes(int P, int nx, int ny, int nz);
size_t r2x_binning_bytes(long long R);             /* shared by rasterizer and voxelizer */
size_t r2x_raster_bwd_scratch_bytes(long long R);  /* per-instance moment buffer of the backward pass */
size_t r2x_voxel_bwd_scratch_bytes(long long R);

/* ---- rasterizer (X-ray projection) --------------------------------------------------------- */
/* Synchronous: one host<->device round trip to learn R (as the reference), binning buffer obtained
 * from `binning_alloc(r2x_binning_bytes(R), alloc_user)`.  *num_rendered (host) receives R. */
int r2x_raster_forward(void* stream, int P, int W, int H, const float* means3D, const float* opacities,
                       const float* scales, float scale_modifier, const float* rotations,
                       const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                       const float* campos, float tan_fovx, float tan_fovy, int prefiltered, int mode,
                       float* out_color, int* radii, void* geom_buf, void* image_buf,
                       r2x_alloc_fn binning_alloc, void* alloc_user, int debug, int* num_rendered);

/* Asynchronous: no host synchronisation.  `binning_buf` must hold r2x_binning_bytes(capacity); if the
 * scene needs more than `capacity` instances the extra ones are dropped and the overflow is reported
 * through `status_dev` (device uint32[2]: {R, overflow flag}), which the caller inspects after it
 * synchronises for its own reasons. */
int r2x_raster_forward_async(void* stream, int P, int W, int H, const float* means3D, const float* opacities,
                             const float* scales, float scale_modifier, const float* rotations,
                             const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                             const float* campos, float tan_fovx, float tan_fovy, int prefiltered, int mode,
                             float* out_color, int* radii, void* geom_buf, void* image_buf, void* binning_buf,
                             long long capacity, uint32_t* status_dev);

/* `R` is the instance count the binning buffer was carved for (num_rendered of the synchronous call,
 * `capacity` of the asynchronous one).  `scratch` holds r2x_raster_bwd_scratch_bytes(R).  All eight
 * gradient arrays are fully written (no pre-zeroing needed): dL_dmean2D[P,3], dL_dopacity[P],
 * dL_dmu[P] (may be NULL), dL_dmean3D[P,3], dL_dcov3D[P,6], dL_dscale[P,3], dL_drot[P,4]. */
int r2x_raster_backward(void* stream, int P, long long R, int W, int H, const float* means3D,
                        const float* scales, float scale_modifier, const float* rotations,
                        const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                        const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                        const void* geom_buf, const void* binning_buf, const void* image_buf, void* scratch,
                        const float* dL_dpix, float* dL_dmean2D, float* dL_dopacity, float* dL_dmu,
                        float* dL_dmean3D, float* dL_dcov3D, float* dL_dscale, float* dL_drot, int mode,
                        int debug);

/* Stage entry points for measurement (bench.py roofline, ncu): re-run ONLY the per-tile accumulation
 * kernel (the reference's renderCUDA, RAS/forward.cu:294-395 / VOX/forward.cu:183-315) on the state a
 * previous forward left in the three buffers.  `R` = the count the binning buffer was carved for. */
int r2x_raster_render_only(void* stream, int P, int W, int H, long long R, const void* geom_buf,
                           const void* binning_buf, const void* image_buf, float* out_color);
int r2x_voxel_render_only(void* stream, int P, int nx, int ny, int nz, long long R, const void* geom_buf,
                          const void* binning_buf, const void* image_buf, float* out_volume);

int r2x_mark_visible(void* stream, int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, unsigned char* present);

/* Stage outputs in the reference's layouts (any pointer may be NULL): means2D[P,2], depths[P],
 * conic_opacity[P,4], mus[P], tiles_touched[P], point_offsets[P], keys[R] = (tile<<32)|depth_bits for
 * each entry of point_list[R] (our sorted order: tile-major, Gaussian index ascending), ranges[T,2]. */
int r2x_raster_export(void* stream, int P, int W, int H, long long R, const void* geom_buf,
                      const void* binning_buf, const void* image_buf, float* means2D, float* depths,
                      float* conic_opacity, float* mus, uint32_t* tiles_touched, uint32_t* point_offsets,
                      uint64_t* keys, uint32_t* point_list, uint32_t* ranges);

/* ---- voxelizer (density volume) ------------------------------------------------------------ */
int r2x_voxel_forward(void* stream, int P, int nx, int ny, int nz, float sx, float sy, float sz, float cx,
                      float cy, float cz, const float* means3D, const float* opacities, const float* scales,
                      float scale_modifier, const float* rotations, const float* cov3D_precomp,
                      int prefiltered, float* out_volume, int* radii_x, int* radii_y, int* radii_z,
                      void* geom_buf, void* image_buf, r2x_alloc_fn binning_alloc, void* alloc_user, int debug,
                      int* num_rendered);

int r2x_voxel_forward_async(void* stream, int P, int nx, int ny, int nz, float sx, float sy, float sz,
                            float cx, float cy, float cz, const float* means3D, const float* opacities,
                            const float* scales, float scale_modifier, const float* rotations,
                            const float* cov3D_precomp, int prefiltered, float* out_volume, int* radii_x,
                            int* radii_y, int* radii_z, void* geom_buf, void* image_buf, void* binning_buf,
                            long long capacity, uint32_t* status_dev);

/* Gradients fully written: dL_dopacity[P], dL_dmean3D[P,3], dL_dcov3D[P,6], dL_dscale[P,3], dL_drot[P,4]. */
int r2x_voxel_backward(void* stream, int P, long long R, int nx, int ny, int nz, float sx, float sy, float sz,
                       float cx, float cy, float cz, const float* means3D, const float* scales,
                       float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const int* radii_x, const int* radii_y, const int* radii_z, const void* geom_buf,
                       const void* binning_buf, const void* image_buf, void* scratch, const float* dL_dvol,
                       float* dL_dopacity, float* dL_dmean3D, float* dL_dcov3D, float* dL_dscale,
                       float* dL_drot, int debug);

/* means3D_norm[P,3], depths[P], conic_opacity[P,7] (a,b,c,d,e,f,rho), others as r2x_raster_export. */
int r2x_voxel_export(void* stream, int P, int nx, int ny, int nz, long long R, const void* geom_buf,
                     const void* binning_buf, const void* image_buf, float* means3D_norm, float* depths,
                     float* conic_opacity, uint32_t* tiles_touched, uint32_t* point_offsets, uint64_t* keys,
                     uint32_t* point_list, uint32_t* ranges);

#ifdef __cplusplus
}
#endif
#endif /* R2X_H_INCLUDED */
