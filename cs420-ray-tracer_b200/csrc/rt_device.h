// rt_device.h -- device-side data layout shared by the kernels and the host API.
//
// Layout in HBM (per rt_ctx, all built by rt_upload_scene):
//   sph64   N x double4  (cx, cy, cz, r*r)         exact geometry, read by the FP64 deciders
//   mat     N x float4   (R, G, B, reflectivity)   shading only
//   matx    N x float2   (shininess, recurse flag) flag = (reflectivity > 0) evaluated in double
//   cam32 / light32[l]   N/2 x {float4,float4}     FP32 filter tables, sphere PAIRS, one table per
//                                                  shared ray origin (camera, each light)
//   sph32   N/2 x {float4 x3}                      FP32 filter table for general-origin rays
// Camera basis, lights and ambient live in __constant__ memory (g_frame).
#ifndef RT_DEVICE_H
#define RT_DEVICE_H

#include <cuda_runtime.h>
#include <stdint.h>

#define RT_MAX_LIGHTS 64

struct RtFrameConst {
  double cam_pos[3];
  double fwd[3], right[3], up[3];   // include/camera.h:10-15, computed on the host in double
  double light_pos[RT_MAX_LIGHTS][3];
  float light_col[RT_MAX_LIGHTS][3];
  float ambient[3];
  int nlights;
  int nspheres;
};

// Row-band mapping: local row lr of rank `rank` -> global row j.
struct RtBands {
  int band_h, rank, nranks, local_rows;
};

__host__ __device__ inline int rt_local_to_global_row(const RtBands &b, int lr) {
  if (b.nranks == 1) return lr;                     // (one rank owns every band: no integer division per pixel)
  int lb = lr / b.band_h;
  return (lb * b.nranks + b.rank) * b.band_h + (lr - lb * b.band_h);
}

// Device counters (uint64 each); indices into the counter array.
enum {
  RT_CNT_CLOSEST = 0,
  RT_CNT_HITS = 1,
  RT_CNT_SHADOW = 2,
  RT_CNT_OCCLUDED = 3,
  RT_CNT_FP64 = 4,
  RT_CNT_TESTS = 5,
  RT_CNT_VIOLATIONS = 6,
  RT_CNT_ALIVE0 = 8,   // .. RT_CNT_ALIVE0 + 31
  RT_CNT_CAND = 40,    // spheres left after bundle culling, summed over warp-level table walks
  RT_CNT_WALKS = 41,   // warp-level table walks that were bundle-culled
  RT_CNT_FALLBACKS = 42,  // LBVH bundles that fell back to one traversal per ray (too wide / frontier overflow)
  RT_CNT_TOTAL = 48
};

struct RtRenderArgs {
  int W, H, max_depth;
  RtBands bands;
  const double *su;        // W: (i/(W-1) - 0.5) * scale * aspect   (include/camera.h:21-22)
  const double *sv;        // H: (j/(H-1) - 0.5) * scale
  const double4 *sph64;
  const float4 *mat;
  const float2 *matx;
  uint8_t *rgb;            // local_rows x W x 3 (8-bit output; unused when fb is set)
  float *fb;               // optional FLOAT output, 3 floats per pixel (tile renders, supersampling): replaces rgb
  int out_remap;           // 0: output pixel = local pixel lr*W + x; 1: (lr + out_y0) * out_pitch + x + out_x0;
                           // 2: global_row(lr) * W + x -- rows land at their image positions of an assembled frame
  int out_pitch, out_x0, out_y0;
  int32_t *hit_idx;        // optional debug, local_rows*W*max_depth
  uint32_t *shadow_mask;   // optional debug
  unsigned long long *counters;  // optional
};

#endif
