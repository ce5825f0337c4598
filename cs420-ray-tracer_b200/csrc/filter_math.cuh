// filter_math.cuh -- rigorous FP32 interval refinement of one (ray, sphere) candidate.
//
// The fast kernels test every ray against every sphere with a 4-FMA (shared ray origin) or
// 10-op (general origin) FP32 expression whose sign is CONSERVATIVE: it may flag a sphere the
// reference misses, never the reverse.  Flagged spheres come here.  From FP32 data and explicit
// error bounds this code brackets the two roots of the reference's quadratic
// (include/sphere.h:29-47) with directed-rounding arithmetic and classifies the sphere as
// MISS / HIT with t in [lo, hi] / AMBIGUOUS.  Only AMBIGUOUS cases (and hits whose brackets
// overlap) are re-evaluated with the exact FP64 routine of exact_fp64.cuh, so every decision
// the reference makes (hit?, which root, which sphere is closest, occluded?) is reproduced
// bit-for-bit while >99 % of the arithmetic stays in FP32.
//
// Error model (u = 2^-24).  Inputs: ocf = fl32(c - O) with c - O formed in double; a float
// direction dt with |dt - d*| <= 12u for the true unit direction d*; tca = fl32 dot(ocf, dt)
// with FMAs.  Then |tca - oc.d*| <= (3u + u + 12u)|oc| = RT_ETA |oc|  (RT_ETA = 2^-20), and
// D* = (oc.d*)^2 - (|oc|^2 - r^2) lies within RT_EK*RT_ETA*|oc|^2 + delta64 of the FP32 value.
#ifndef RT_FILTER_MATH_CUH
#define RT_FILTER_MATH_CUH

#include <cuda_runtime.h>

#define RT_ETA 9.5367431640625e-07f   /* 2^-20 */
#define RT_EK 2.02f

namespace rtf {

struct Roots {
  float n_lo, n_hi;   // bracket of the smaller root (t1 of include/sphere.h:47)
  float f_lo, f_hi;   // bracket of the larger root  (t2 of include/sphere.h:48)
};

enum { RT_MISS = 0, RT_HIT = 1, RT_AMBIG = 2 };

// Half-width of the discriminant's uncertainty for a sphere at squared distance oc2 from the ray
// origin (or from the shared origin O); d64 = absolute FP64/geometry slack of the scene.
__device__ __forceinline__ float disc_margin(float oc2, float d64) {
  return __fmaf_ru(RT_EK * RT_ETA, oc2, d64);
}

// Brackets both roots given tca (distance along the ray to the closest approach), Dhi = an UPPER
// bound of the true quarter-discriminant D*, E2 = width such that D* >= Dhi - E2, and dt = bound of
// |tca - true|.  Returns false when the sign of D* is not certain (tangent zone) -> AMBIGUOUS.
__device__ __forceinline__ bool bracket_roots(float tca, float Dhi, float E2, float dt, Roots &r) {
  float Dlo = __fsub_rd(Dhi, E2);
  if (!(Dlo > 0.0f)) return false;
  float s_lo = __fsqrt_rd(Dlo), s_hi = __fsqrt_ru(Dhi);
  float c_lo = __fsub_rd(tca, dt), c_hi = __fadd_ru(tca, dt);
  r.n_lo = __fsub_rd(c_lo, s_hi); r.n_hi = __fsub_ru(c_hi, s_lo);
  r.f_lo = __fadd_rd(c_lo, s_lo); r.f_hi = __fadd_ru(c_hi, s_hi);
  return true;
}

// include/sphere.h:49-56 on brackets: which root does the reference return?
//   max(t1,t2) < 0 -> miss;  t = min; if (t < 0) t = max.
__device__ __forceinline__ int select_root(const Roots &r, float &lo, float &hi) {
  if (r.f_hi < 0.0f) return RT_MISS;                                   // both roots negative
  if (r.n_lo > 0.0f) { lo = r.n_lo; hi = r.n_hi; return RT_HIT; }      // near root is the answer
  if (r.n_hi < 0.0f && r.f_lo > 0.0f) { lo = r.f_lo; hi = r.f_hi; return RT_HIT; }   // origin inside
  return RT_AMBIG;
}

}  // namespace rtf
#endif
