// kernels_exact.cuh -- "mode 1": the whole hot path in FP64, brute force, one thread per pixel.
//
// This is the on-device statement of the reference semantics (src/main.cpp:16-58,146-157 and
// include/scene.h:41-121) with nothing clever in it: every ray tests every sphere with the
// exact routine of exact_fp64.cuh, shading is FP64 too.  It is the diagnostic mode of the
// library (rt_set_option "mode"=1) and the device-side cross-check for the fast kernels at
// sizes where the CPU oracle would take hours; it is not the performance path.
#ifndef RT_KERNELS_EXACT_CUH
#define RT_KERNELS_EXACT_CUH

#include "exact_fp64.cuh"
#include "rt_device.h"

namespace rtk {

using rtx::d3;

__device__ __forceinline__ d3 ld3(const double *p) { return rtx::mk(p[0], p[1], p[2]); }
// (cx, cy, cz, r*r) through the read-only path as two 16-byte loads
__device__ __forceinline__ double4 ld_sph(const double4 *p) {
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// include/scene.h:41-61 -- strict '<' in ascending index order
__device__ __forceinline__ int exact_closest(const double4 *__restrict__ sph, int n, d3 o, d3 d, double &tbest) {
  double a = rtx::dot(d, d);
  double t = 1e20;   // INFINITY_DOUBLE, include/ray_math_constants.h:23
  int idx = -1;
  for (int i = 0; i < n; i++) {
    double4 s = ld_sph(&sph[i]);
    double tt;
    if (rtx::intersect(o, d, a, rtx::mk(s.x, s.y, s.z), s.w, tt)) {
      if (tt < t) { t = tt; idx = i; }
    }
  }
  tbest = t;
  return idx;
}

// include/scene.h:65-86 -- "closest t < light distance" == "any sphere with t < light distance"
__device__ __forceinline__ bool exact_occluded(const double4 *__restrict__ sph, int n, d3 p, d3 lpos) {
  d3 o, d; double ldist;
  rtx::shadow_ray(p, lpos, 0.001, o, d, ldist);
  double a = rtx::dot(d, d);
  for (int i = 0; i < n; i++) {
    double4 s = ld_sph(&sph[i]);
    double tt;
    if (rtx::intersect(o, d, a, rtx::mk(s.x, s.y, s.z), s.w, tt)) {
      if (tt < 1e20 && tt < ldist) return true;
    }
  }
  return false;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// src/main.cpp:84-86
__device__ __forceinline__ unsigned quantise(double c) { return (unsigned)(int)(255.99 * (c < 1.0 ? c : 1.0)); }

__global__ void __launch_bounds__(128) k_exact(RtRenderArgs a) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int lr = blockIdx.y * 4 + (threadIdx.x >> 5);
  const bool live = x < a.W && lr < a.bands.local_rows;
  unsigned long long c_closest = 0, c_hits = 0, c_shadow = 0, c_occ = 0, c_tests = 0;
  const int n = g_frame.nspheres, nl = g_frame.nlights;
  if (live) {
    const int j = rt_local_to_global_row(a.bands, lr);
    const size_t p = (size_t)lr * a.W + x;
    int32_t *hi = a.hit_idx ? a.hit_idx + p * a.max_depth : nullptr;
    uint32_t *sm = a.shadow_mask ? a.shadow_mask + p * a.max_depth : nullptr;
    if (hi) for (int k = 0; k < a.max_depth; k++) hi[k] = -2;
    if (sm) for (int k = 0; k < a.max_depth; k++) sm[k] = 0;
    d3 o = ld3(g_frame.cam_pos);
    d3 d = rtx::camera_dir(ld3(g_frame.fwd), ld3(g_frame.right), ld3(g_frame.up), a.su[x], a.sv[j]);
    double wt = 1.0, ar = 0, ag = 0, ab = 0;
    for (int level = 0; level < a.max_depth; level++) {
      double t;
      c_closest++; c_tests += n;
      if (a.counters && level < 32) atomicAdd(&a.counters[RT_CNT_ALIVE0 + level], 1ull);
      int idx = exact_closest(a.sph64, n, o, d, t);
      if (hi) hi[level] = idx;
      if (idx < 0) {  // src/main.cpp:26-30
        double ts = 0.5 * (d.y + 1.0);
        ar += wt * (1.0 * (1.0 - ts) + 0.5 * ts);
        ag += wt * (1.0 * (1.0 - ts) + 0.7 * ts);
        ab += wt * (1.0 * (1.0 - ts) + 1.0 * ts);
        break;
      }
      c_hits++;
      double4 s = ld_sph(&a.sph64[idx]);
      float4 m = __ldg(&a.mat[idx]);
      float2 mx = __ldg(&a.matx[idx]);
      d3 hit = rtx::hit_point(o, d, t);
      d3 nrm = rtx::normal_at(hit, rtx::mk(s.x, s.y, s.z));
      d3 view = rtx::normalized(rtx::sub(o, hit));
      // include/scene.h:89-121
      double cr = (double)g_frame.ambient[0] * m.x, cg = (double)g_frame.ambient[1] * m.y, cb = (double)g_frame.ambient[2] * m.z;
      uint32_t mask = 0;
      for (int l = 0; l < nl; l++) {
        d3 lp = ld3(g_frame.light_pos[l]);
        c_shadow++; c_tests += n;
        if (exact_occluded(a.sph64, n, hit, lp)) { c_occ++; if (l < 32) mask |= 1u << l; continue; }
        d3 ldir = rtx::normalized(rtx::sub(lp, hit));
        double ndl = fmax(0.0, rtx::dot(nrm, ldir));
        double kd = (1.0 - (double)m.w) * ndl;
        d3 nl2 = rtx::scale(ldir, -1.0);
        d3 rdir = rtx::sub(nl2, rtx::scale(rtx::scale(nrm, 2.0), rtx::dot(nl2, nrm)));
        double rdv = fmax(0.0, rtx::dot(rdir, view));
        double spec = 0.5 * pow(rdv, (double)mx.x);
        cr += g_frame.light_col[l][0] * spec + m.x * kd;
        cg += g_frame.light_col[l][1] * spec + m.y * kd;
        cb += g_frame.light_col[l][2] * spec + m.z * kd;
      }
      if (sm) sm[level] = mask;
      if (mx.y > 0.5f) {  // reflectivity > 0 (src/main.cpp:43); blend :53-54 unrolled front to back
        double refl = (double)m.w;
        ar += wt * (1.0 - refl) * cr; ag += wt * (1.0 - refl) * cg; ab += wt * (1.0 - refl) * cb;
        wt *= refl;
        d3 o2, d2;
        rtx::reflect_ray(d, hit, nrm, 0.001, o2, d2);
        o = o2; d = d2;
      } else {
        ar += wt * cr; ag += wt * cg; ab += wt * cb;
        break;
      }
    }
    uint8_t *px = a.rgb + p * 3;
    px[0] = (uint8_t)quantise(ar); px[1] = (uint8_t)quantise(ag); px[2] = (uint8_t)quantise(ab);
  }
  if (a.counters) {
    c_closest = warp_sum(c_closest); c_hits = warp_sum(c_hits); c_shadow = warp_sum(c_shadow);
    c_occ = warp_sum(c_occ); c_tests = warp_sum(c_tests);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&a.counters[RT_CNT_CLOSEST], c_closest);
      atomicAdd(&a.counters[RT_CNT_HITS], c_hits);
      atomicAdd(&a.counters[RT_CNT_SHADOW], c_shadow);
      atomicAdd(&a.counters[RT_CNT_OCCLUDED], c_occ);
      atomicAdd(&a.counters[RT_CNT_TESTS], c_tests);
      atomicAdd(&a.counters[RT_CNT_FP64], c_tests);
    }
  }
}

}  // namespace rtk
#endif
