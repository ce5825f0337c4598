// kernels_fast.cuh -- "mode 0": the production path.  sm_100a only.
//
// Work decomposition (wavefront with compacted queues):
//   k_primary   one persistent CTA per SM slot; tiles of 32x16 pixels, two pixels per thread.
//               camera rays -> closest hit (FP32 filter over the CAMERA table, FP64 decide)
//               -> Phong + one any-hit shadow query per light (FP32 filter over that LIGHT's
//               table) -> final 8-bit pixel, or a reflected ray appended to the ray queue
//               (warp-ballot compaction: only live rays reach the next level).
//   k_bounce    one launch per reflection level >= 1: consumes the queue (two rays per thread),
//               general-origin closest hit, same shading, appends to the other queue.
// Sphere tables are staged once per CTA into shared memory with one TMA bulk copy
// (cp.async.bulk + mbarrier) when they fit, otherwise read through L1/L2.
//
// What is FP32 and what is FP64:  every ray/sphere TEST is 4 packed-FP32 FMAs per sphere pair
// (FFMA2; shared origin) or 10 (general origin).  The tests are conservative (filter_math.cuh);
// the few spheres they flag are bracketed in FP32 interval arithmetic and, only where brackets
// touch, decided by the reference's own FP64 formula (exact_fp64.cuh).  Hit points, normals and
// reflected rays -- everything that feeds the NEXT query -- are FP64 in the reference's
// operation order, so hit indices and shadow booleans are bit-exact.  Colour is FP32.
#ifndef RT_KERNELS_FAST_CUH
#define RT_KERNELS_FAST_CUH

#include "exact_fp64.cuh"
#include "filter_math.cuh"
#include "rt_device.h"

namespace rtf {

using rtx::d3;

constexpr int kThreads = 256;
constexpr int kTileW = 32, kTileH = 16;
constexpr int kGroupPairs = 4;          // sphere pairs per fast-path group (8 spheres)
constexpr int kSmemHeader = 2048;       // mbarrier, tile index, RGB staging
constexpr float kEps = 0.001f;          // EPSILON, include/ray_math_constants.h:22

struct __align__(16) RayRec {           // one queued reflected ray (80 bytes)
  double ox, oy, oz, dx, dy, dz;        // exact FP64 origin / unit direction (src/main.cpp:45-48)
  unsigned pix;                         // local pixel index lr*W + x
  float wt;                             // product of reflectivities so far
  float ar, ag, ab;                     // colour accumulated so far (front to back)
  unsigned pad;
};

struct FastArgs {
  RtRenderArgs r;
  const float4 *otab;     // (1+L) tables x npairs x 2 float4; table 0 = camera, 1+l = light l
  const float4 *gtab;     // general-origin table: npairs x 2 float4 (recentred centres, rho')
  int npairs;             // padded to a multiple of kGroupPairs
  int N, L;
  float d64;              // absolute slack covering FP64 rounding / geometry (delta64)
  float gS2;              // squared radius bound of the recentred scene (general filter)
  double c0[3];           // recentring offset of the general table
  int tiles_x, ntiles;
  unsigned int *tile_counter;
  RayRec *q_out; unsigned int *q_out_count;
  const RayRec *q_in; const unsigned int *q_in_count;
  unsigned int *chunk_counter;
  int level;
  int tables_in_smem;
  unsigned table_bytes;   // bytes staged into shared memory
};

// ---------------------------------------------------------------------------------------------
// small helpers
__device__ __forceinline__ unsigned fbits(float x) { return __float_as_uint(x); }
__device__ __forceinline__ float f_rd(double x) { return __double2float_rd(x); }
__device__ __forceinline__ float f_ru(double x) { return __double2float_ru(x); }
__device__ __forceinline__ d3 ldc3(const double *p) { return rtx::mk(p[0], p[1], p[2]); }
__device__ __forceinline__ double4 ld_sph64(const double4 *p) {
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ unsigned long long wsum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned quant8(float c) { return (unsigned)(int)(255.99f * fminf(1.0f, c)); }

// mbarrier + TMA bulk copy (global -> shared), one phase, used once per CTA
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src),
               "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(a), "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// exact (FP64) deciders, kept out of line: they are rare and register hungry
struct ExactRay { d3 o, d; double a; };

__device__ __noinline__ ExactRay exact_primary_ray(const double *su, const double *sv, int x, int j) {
  ExactRay e;
  e.o = ldc3(g_frame.cam_pos);
  e.d = rtx::camera_dir(ldc3(g_frame.fwd), ldc3(g_frame.right), ldc3(g_frame.up), su[x], sv[j]);
  e.a = rtx::dot(e.d, e.d);
  return e;
}
__device__ __noinline__ ExactRay exact_shadow_ray(d3 p, int light, double &ldist) {
  ExactRay e;
  rtx::shadow_ray(p, ldc3(g_frame.light_pos[light]), 0.001, e.o, e.d, ldist);
  e.a = rtx::dot(e.d, e.d);
  return e;
}
__device__ __noinline__ bool exact_sphere(const double4 *sph64, int idx, d3 o, d3 d, double a, double &t) {
  double4 s = ld_sph64(&sph64[idx]);
  return rtx::intersect(o, d, a, rtx::mk(s.x, s.y, s.z), s.w, t);
}
// safety net only: full FP64 brute force for one ray (a filter violation was detected)
__device__ __noinline__ int exact_bruteforce(const double4 *sph64, int n, d3 o, d3 d, double a, double &tbest) {
  double t = 1e20; int idx = -1;
  for (int i = 0; i < n; i++) {
    double tt;
    if (exact_sphere(sph64, i, o, d, a, tt) && tt < t) { t = tt; idx = i; }
  }
  tbest = t;
  return idx;
}

// ---------------------------------------------------------------------------------------------
// closest-hit bookkeeping: best candidate as an FP32 bracket, exact FP64 t only when needed
struct Best {
  float lo, hi;
  int idx;
  double t;       // valid iff exact
  bool exact;
};
__device__ __forceinline__ void best_init(Best &b) { b.lo = 3.0e38f; b.hi = 3.0e38f; b.idx = -1; b.t = 1e20; b.exact = false; }
__device__ __forceinline__ void best_set_exact(Best &b, int idx, double t) {
  b.idx = idx; b.t = t; b.exact = true; b.lo = f_rd(t); b.hi = f_ru(t);
}

// One candidate sphere `i` of a closest-hit query.  status/lo/hi from the FP32 brackets; `ray`
// is only materialised (by the caller-supplied functor) when an FP64 decision is unavoidable.
template <typename ExactRayFn>
__device__ __forceinline__ void closest_consider(Best &b, int i, int status, float lo, float hi, const double4 *sph64,
                                                 ExactRayFn get_ray, unsigned &n_fp64) {
  if (status == RT_MISS) return;
  if (status == RT_HIT) {
    if (b.idx < 0) { b.lo = lo; b.hi = hi; b.idx = i; b.exact = false; return; }
    if (lo >= b.hi) return;                       // cannot be strictly closer (include/scene.h:52)
    if (hi < b.lo) { b.lo = lo; b.hi = hi; b.idx = i; b.exact = false; return; }
  }
  // brackets overlap, or the sphere itself is ambiguous: decide in FP64, lowest index wins ties
  ExactRay e = get_ray();
  double tn;
  n_fp64++;
  bool hn = exact_sphere(sph64, i, e.o, e.d, e.a, tn);
  if (!hn || !(tn < 1e20)) return;                // INFINITY_DOUBLE init of include/scene.h:42
  if (b.idx >= 0 && !b.exact) {
    float tl = f_rd(tn), th = f_ru(tn);
    if (tl >= b.hi) return;
    if (th < b.lo) { best_set_exact(b, i, tn); return; }
    double tb;
    n_fp64++;
    bool hb = exact_sphere(sph64, b.idx, e.o, e.d, e.a, tb);
    if (hb) best_set_exact(b, b.idx, tb); else best_init(b);   // (else: filter violation, caught later)
  }
  if (b.idx < 0 || tn < b.t || (tn == b.t && i < b.idx)) best_set_exact(b, i, tn);
}

// ---------------------------------------------------------------------------------------------
// table access: sphere pair p of a table = two float4: (x0,x1,y0,y1) (z0,z1,w0,w1)
template <bool kSmem>
__device__ __forceinline__ float4 tab_ld(const float4 *t, int i) {
  if (kSmem) return t[i];
  return __ldg(&t[i]);
}

// scalar re-evaluation of one sphere of a shared-origin table (same operations as the fast path)
__device__ __forceinline__ void shared_origin_eval(float ocx, float ocy, float ocz, float ncc, float dx, float dy, float dz,
                                                   float &tca, float &Dp) {
  tca = __fmul_rn(ocx, dx);
  tca = __fmaf_rn(ocy, dy, tca);
  tca = __fmaf_rn(ocz, dz, tca);
  Dp = __fmaf_rn(tca, tca, ncc);
}

// Brackets the roots of sphere (oc, ncc) of a shared-origin table for direction d.
// Returns false if the discriminant's sign is uncertain.
__device__ __forceinline__ bool shared_origin_roots(float ocx, float ocy, float ocz, float ncc, float tca, float Dp, float d64,
                                                    Roots &r) {
  float oc2 = __fmaf_ru(ocz, ocz, __fmaf_ru(ocy, ocy, __fmul_ru(ocx, ocx)));
  float E = disc_margin(oc2, d64);
  // D* <= Dp (+ rounding), D* >= Dp - 2E - ulp(ncc) (- rounding): see DESIGN.md "filter margins"
  float slop = __fmul_ru(4.8e-7f, fabsf(Dp) + fabsf(ncc));
  float Dhi = __fadd_ru(Dp, slop);
  float E2 = __fadd_ru(__fadd_ru(__fmul_ru(2.0f, E), slop), slop);
  float dt = __fmul_ru(RT_ETA * 1.001f, __fsqrt_ru(oc2));
  return bracket_roots(tca, Dhi, E2, dt, r);
}

// ---------------------------------------------------------------------------------------------
// CLOSEST HIT, shared origin (camera table).  Two rays per thread.
template <bool kSmem, typename ExactRayFn0, typename ExactRayFn1>
__device__ __forceinline__ void closest_shared(const float4 *__restrict__ tab, int npairs, int N, const float (&dx)[2],
                                               const float (&dy)[2], const float (&dz)[2], const bool (&live)[2], float d64,
                                               const double4 *sph64, ExactRayFn0 ray0, ExactRayFn1 ray1, Best (&best)[2],
                                               unsigned &n_fp64) {
  const unsigned dead0 = live[0] ? 0u : 0x80000000u, dead1 = live[1] ? 0u : 0x80000000u;
  const float2 dx0 = make_float2(dx[0], dx[0]), dy0 = make_float2(dy[0], dy[0]), dz0 = make_float2(dz[0], dz[0]);
  const float2 dx1 = make_float2(dx[1], dx[1]), dy1 = make_float2(dy[1], dy[1]), dz1 = make_float2(dz[1], dz[1]);
  for (int g = 0; g < npairs; g += kGroupPairs) {
    unsigned acc0 = 0xffffffffu, acc1 = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < kGroupPairs; k++) {
      const float4 A = tab_ld<kSmem>(tab, 2 * (g + k)), B = tab_ld<kSmem>(tab, 2 * (g + k) + 1);
      const float2 X = make_float2(A.x, A.y), Y = make_float2(A.z, A.w), Z = make_float2(B.x, B.y), W = make_float2(B.z, B.w);
      float2 t0 = __fmul2_rn(X, dx0); t0 = __ffma2_rn(Y, dy0, t0); t0 = __ffma2_rn(Z, dz0, t0);
      float2 t1 = __fmul2_rn(X, dx1); t1 = __ffma2_rn(Y, dy1, t1); t1 = __ffma2_rn(Z, dz1, t1);
      float2 D0 = __ffma2_rn(t0, t0, W), D1 = __ffma2_rn(t1, t1, W);
      acc0 &= fbits(D0.x) & fbits(D0.y);
      acc1 &= fbits(D1.x) & fbits(D1.y);
    }
    if ((int)((acc0 | dead0) & (acc1 | dead1)) >= 0) {
      // slow path: some sphere of this group may be hit by one of this thread's rays
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if ((int)((r ? acc1 : acc0) | (r ? dead1 : dead0)) < 0) continue;
        for (int k = 0; k < 2 * kGroupPairs; k++) {
          const int pi = g + (k >> 1), h = k & 1, i = 2 * pi + h;
          if (i >= N) break;
          const float4 A = tab_ld<kSmem>(tab, 2 * pi), B = tab_ld<kSmem>(tab, 2 * pi + 1);
          const float ocx = h ? A.y : A.x, ocy = h ? A.w : A.z, ocz = h ? B.y : B.x, ncc = h ? B.w : B.z;
          float tca, Dp;
          shared_origin_eval(ocx, ocy, ocz, ncc, dx[r], dy[r], dz[r], tca, Dp);
          if (!(Dp >= 0.0f)) continue;
          Roots rt;
          int status = RT_AMBIG;
          float lo = 0, hi = 0;
          if (shared_origin_roots(ocx, ocy, ocz, ncc, tca, Dp, d64, rt)) status = select_root(rt, lo, hi);
          if (r == 0) closest_consider(best[0], i, status, lo, hi, sph64, ray0, n_fp64);
          else closest_consider(best[1], i, status, lo, hi, sph64, ray1, n_fp64);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// CLOSEST HIT, general origin (bounce rays).  Table pair = (cx0,cx1,cy0,cy1) (cz0,cz1,rho0,rho1)
// with recentred centres; o = recentred float origin, di = direction inflated by (1+16u).
template <bool kSmem, typename ExactRayFn0, typename ExactRayFn1>
__device__ __forceinline__ void closest_general(const float4 *__restrict__ tab, int npairs, int N, const float (&ox)[2],
                                                const float (&oy)[2], const float (&oz)[2], const float (&dx)[2],
                                                const float (&dy)[2], const float (&dz)[2], const bool (&live)[2], float d64,
                                                float gS2, const double4 *sph64, ExactRayFn0 ray0, ExactRayFn1 ray1,
                                                Best (&best)[2], unsigned &n_fp64) {
  const unsigned dead0 = live[0] ? 0u : 0x80000000u, dead1 = live[1] ? 0u : 0x80000000u;
  const float kInfl = 1.0f + 16.0f * 5.9604645e-8f;
  float2 nox[2], noy[2], noz[2], idx2[2], idy2[2], idz2[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    nox[r] = make_float2(-ox[r], -ox[r]); noy[r] = make_float2(-oy[r], -oy[r]); noz[r] = make_float2(-oz[r], -oz[r]);
    float ix = __fmul_rn(dx[r], kInfl), iy = __fmul_rn(dy[r], kInfl), iz = __fmul_rn(dz[r], kInfl);
    idx2[r] = make_float2(ix, ix); idy2[r] = make_float2(iy, iy); idz2[r] = make_float2(iz, iz);
  }
  for (int g = 0; g < npairs; g += kGroupPairs) {
    unsigned acc[2] = {0xffffffffu, 0xffffffffu};
#pragma unroll
    for (int k = 0; k < kGroupPairs; k++) {
      const float4 A = tab_ld<kSmem>(tab, 2 * (g + k)), B = tab_ld<kSmem>(tab, 2 * (g + k) + 1);
      const float2 CX = make_float2(A.x, A.y), CY = make_float2(A.z, A.w), CZ = make_float2(B.x, B.y);
      const float2 NR = make_float2(-B.z, -B.w);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        float2 x = __fadd2_rn(CX, nox[r]), y = __fadd2_rn(CY, noy[r]), z = __fadd2_rn(CZ, noz[r]);
        float2 t = __fmul2_rn(x, idx2[r]); t = __ffma2_rn(y, idy2[r], t); t = __ffma2_rn(z, idz2[r], t);
        float2 q = __ffma2_rn(x, x, NR); q = __ffma2_rn(y, y, q); q = __ffma2_rn(z, z, q);
        float2 D = __ffma2_rn(t, t, make_float2(-q.x, -q.y));
        acc[r] &= fbits(D.x) & fbits(D.y);
      }
    }
    if ((int)((acc[0] | dead0) & (acc[1] | dead1)) >= 0) {
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if ((int)(acc[r] | (r ? dead1 : dead0)) < 0) continue;
        for (int k = 0; k < 2 * kGroupPairs; k++) {
          const int pi = g + (k >> 1), h = k & 1, i = 2 * pi + h;
          if (i >= N) break;
          const float4 A = tab_ld<kSmem>(tab, 2 * pi), B = tab_ld<kSmem>(tab, 2 * pi + 1);
          const float cx = h ? A.y : A.x, cy = h ? A.w : A.z, cz = h ? B.y : B.x, rho = h ? B.w : B.z;
          // explicit-margin evaluation (independent of the inflation trick of the fast path)
          const float x = __fsub_rn(cx, ox[r]), y = __fsub_rn(cy, oy[r]), z = __fsub_rn(cz, oz[r]);
          float tca = __fmul_rn(x, dx[r]); tca = __fmaf_rn(y, dy[r], tca); tca = __fmaf_rn(z, dz[r], tca);
          const float oc2 = __fmaf_ru(z, z, __fmaf_ru(y, y, __fmul_ru(x, x)));
          // |X - oc*| <= u(|c|+|o|+|oc|); D error <= 20u|oc|^2 + 4.1u S^2 + 3u r^2 + d64  
          const float Eg = __fadd_ru(__fmul_ru(1.9073486e-6f, __fadd_ru(__fadd_ru(oc2, gS2), fabsf(rho))), d64);
          const float Dc = __fmaf_rn(tca, tca, __fsub_rn(rho, oc2));      // rho ~ r^2 (+margins, harmless: covered by Eg)
          const float Dhi = __fadd_ru(Dc, Eg);
          if (!(Dhi >= 0.0f)) continue;
          Roots rt;
          int status = RT_AMBIG;
          float lo = 0, hi = 0;
          const float dt = __fmul_ru(RT_ETA * 1.001f, __fadd_ru(__fsqrt_ru(oc2), __fsqrt_ru(gS2)));
          if (bracket_roots(tca, Dhi, __fmul_ru(3.0f, Eg), dt, rt)) status = select_root(rt, lo, hi);
          if (r == 0) closest_consider(best[0], i, status, lo, hi, sph64, ray0, n_fp64);
          else closest_consider(best[1], i, status, lo, hi, sph64, ray1, n_fp64);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// SHADOW (any hit, early out), light table.  Direction dl = unit vector FROM THE LIGHT TOWARDS the
// shaded point, so = distance light -> shadow-ray origin (= |L-p| - EPS), both FP32 with known error.
// Roots s are measured from the light; the reference's t = so - s (include/scene.h:70-85).
//   occluded  <=>  (-EPS < s2 <= so)  or  (s2 > so and -EPS < s1 <= so)
template <bool kSmem>
__device__ __forceinline__ void shadow_light(const float4 *__restrict__ tab, int npairs, int N, int light,
                                             const float (&dx)[2], const float (&dy)[2], const float (&dz)[2],
                                             const float (&so)[2], const bool (&want)[2], const d3 (&p64)[2], float d64,
                                             const double4 *sph64, bool (&occ)[2], unsigned &n_fp64) {
  unsigned dead[2] = {want[0] ? 0u : 0x80000000u, want[1] ? 0u : 0x80000000u};
  occ[0] = occ[1] = false;
  const float2 dx0 = make_float2(dx[0], dx[0]), dy0 = make_float2(dy[0], dy[0]), dz0 = make_float2(dz[0], dz[0]);
  const float2 dx1 = make_float2(dx[1], dx[1]), dy1 = make_float2(dy[1], dy[1]), dz1 = make_float2(dz[1], dz[1]);
  for (int g = 0; g < npairs; g += kGroupPairs) {
    unsigned acc0 = 0xffffffffu, acc1 = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < kGroupPairs; k++) {
      const float4 A = tab_ld<kSmem>(tab, 2 * (g + k)), B = tab_ld<kSmem>(tab, 2 * (g + k) + 1);
      const float2 X = make_float2(A.x, A.y), Y = make_float2(A.z, A.w), Z = make_float2(B.x, B.y), W = make_float2(B.z, B.w);
      float2 t0 = __fmul2_rn(X, dx0); t0 = __ffma2_rn(Y, dy0, t0); t0 = __ffma2_rn(Z, dz0, t0);
      float2 t1 = __fmul2_rn(X, dx1); t1 = __ffma2_rn(Y, dy1, t1); t1 = __ffma2_rn(Z, dz1, t1);
      float2 D0 = __ffma2_rn(t0, t0, W), D1 = __ffma2_rn(t1, t1, W);
      acc0 &= fbits(D0.x) & fbits(D0.y);
      acc1 &= fbits(D1.x) & fbits(D1.y);
    }
    const bool flagged = (int)((acc0 | dead[0]) & (acc1 | dead[1])) >= 0;
    if (__any_sync(0xffffffffu, flagged)) {          // warp-uniform: the vote below needs every lane
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (!flagged || (int)((r ? acc1 : acc0) | dead[r]) < 0) continue;
        const float m = __fmaf_ru(1.9073486e-6f, so[r] + kEps, 1e-7f);      // 2^-19 * |L-p|: covers the FP32 length error
        const float so_lo = so[r] - m, so_hi = so[r] + m, e_lo = -kEps - m, e_hi = -kEps + m;
        for (int k = 0; k < 2 * kGroupPairs && !occ[r]; k++) {
          const int pi = g + (k >> 1), h = k & 1, i = 2 * pi + h;
          if (i >= N) break;
          const float4 A = tab_ld<kSmem>(tab, 2 * pi), B = tab_ld<kSmem>(tab, 2 * pi + 1);
          const float ocx = h ? A.y : A.x, ocy = h ? A.w : A.z, ocz = h ? B.y : B.x, ncc = h ? B.w : B.z;
          float tca, Dp;
          shared_origin_eval(ocx, ocy, ocz, ncc, dx[r], dy[r], dz[r], tca, Dp);
          if (!(Dp >= 0.0f)) continue;
          Roots rt;
          bool decided = false;
          if (shared_origin_roots(ocx, ocy, ocz, ncc, tca, Dp, d64, rt)) {
            const bool no = (rt.n_lo > so_hi) || (rt.f_hi < e_lo) || (rt.n_hi < e_lo && rt.f_lo > so_hi);
            const bool yes = (rt.f_lo > e_hi && rt.f_hi < so_lo) ||
                             (rt.f_lo > so_hi && rt.n_lo > e_hi && rt.n_hi < so_lo);
            if (no) decided = true;
            else if (yes) { decided = true; occ[r] = true; }
          }
          if (!decided) {       // the reference's own formula on the reference's own shadow ray
            double ldist, tt;
            ExactRay e = exact_shadow_ray(p64[r], light, ldist);
            n_fp64++;
            if (exact_sphere(sph64, i, e.o, e.d, e.a, tt) && tt < 1e20 && tt < ldist) occ[r] = true;
          }
        }
        if (occ[r]) dead[r] = 0x80000000u;
      }
      if (__all_sync(0xffffffffu, (dead[0] & dead[1]) != 0u)) break;     // every ray of the warp is done
    }
  }
}

// ---------------------------------------------------------------------------------------------
// warp-ballot compaction: rays still alive are appended densely to the next level's queue
__device__ __forceinline__ void queue_push(bool want, const RayRec &rec, RayRec *q, unsigned int *count) {
  const unsigned m = __ballot_sync(0xffffffffu, want);
  if (m == 0) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (want) q[base + __popc(m & ((1u << lane) - 1u))] = rec;
}

struct Counters {
  unsigned long long closest, hits, shadow, occluded, fp64, violations;
};

// ---------------------------------------------------------------------------------------------
// Shading of up to two hits per thread + continuation (shared by both kernels).
//   in : hit[r], sphere index, exact FP64 ray (o, d) and t of the hit, float view = -d, carried acc/wt
//   out: final[r] (pixel finished, colour in fr/fg/fb) or a pushed reflected ray.
template <bool kSmem>
__device__ __forceinline__ void shade_and_continue(const FastArgs &a, const float4 *__restrict__ ltabs, const bool (&hit)[2],
                                                   const int (&idx)[2], const d3 (&o64)[2], const d3 (&d64v)[2],
                                                   const double (&t64)[2], const unsigned (&pix)[2], float (&wt)[2],
                                                   float (&cr)[2], float (&cg)[2], float (&cb)[2], bool (&final_)[2],
                                                   int level, Counters &cnt, unsigned &n_fp64) {
  d3 p[2], n[2];
  float nx[2], ny[2], nz[2], vx[2], vy[2], vz[2];
  float4 m[2]; float2 mx[2];
  float sr[2], sg[2], sb[2];
  unsigned smask[2] = {0u, 0u};
#pragma unroll
  for (int r = 0; r < 2; r++) {
    p[r] = rtx::mk(0, 0, 0); n[r] = p[r];
    m[r] = make_float4(0, 0, 0, 0); mx[r] = make_float2(0, 0);
    nx[r] = ny[r] = nz[r] = vx[r] = vy[r] = vz[r] = 0.f; sr[r] = sg[r] = sb[r] = 0.f;
    if (hit[r]) {
      const double4 s = ld_sph64(&a.r.sph64[idx[r]]);
      p[r] = rtx::hit_point(o64[r], d64v[r], t64[r]);                     // src/main.cpp:32
      n[r] = rtx::normal_at(p[r], rtx::mk(s.x, s.y, s.z));               // src/main.cpp:35
      m[r] = __ldg(&a.r.mat[idx[r]]); mx[r] = __ldg(&a.r.matx[idx[r]]);
      nx[r] = (float)n[r].x; ny[r] = (float)n[r].y; nz[r] = (float)n[r].z;
      // view_dir = normalized(origin - hit) = -d up to rounding (src/main.cpp:38); colour only
      vx[r] = -(float)d64v[r].x; vy[r] = -(float)d64v[r].y; vz[r] = -(float)d64v[r].z;
      sr[r] = g_frame.ambient[0] * m[r].x; sg[r] = g_frame.ambient[1] * m[r].y; sb[r] = g_frame.ambient[2] * m[r].z;
    }
  }
  const int L = a.L;
  if (__any_sync(0xffffffffu, hit[0] || hit[1])) {
    for (int l = 0; l < L; l++) {
      float dx[2], dy[2], dz[2], so[2];
      bool occ[2];
      const d3 lp = ldc3(g_frame.light_pos[l]);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        dx[r] = dy[r] = dz[r] = 0.f; so[r] = 0.f;
        if (hit[r]) {
          // direction light -> point: FP64 difference, FP32 normalisation (error <= 12u, see filter_math.cuh)
          const d3 w = rtx::sub(p[r], lp);
          const float wx = (float)w.x, wy = (float)w.y, wz = (float)w.z;
          const float l2 = fmaf(wz, wz, fmaf(wy, wy, wx * wx));
          const float inv = rsqrtf(l2);
          dx[r] = wx * inv; dy[r] = wy * inv; dz[r] = wz * inv;
          so[r] = l2 * inv - kEps;
        }
      }
      shadow_light<kSmem>(ltabs + (size_t)l * a.npairs * 2, a.npairs, a.N, l, dx, dy, dz, so, hit, p, a.d64, a.r.sph64, occ,
                          n_fp64);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (!hit[r]) continue;
        cnt.shadow++;
        if (occ[r]) { cnt.occluded++; if (l < 32) smask[r] |= 1u << l; continue; }
        // include/scene.h:104-117 in FP32; light_dir = -(dx,dy,dz)
        const float ndl = fmaxf(0.0f, -(nx[r] * dx[r] + ny[r] * dy[r] + nz[r] * dz[r]));
        const float kd = (1.0f - m[r].w) * ndl;
        const float dn = dx[r] * nx[r] + dy[r] * ny[r] + dz[r] * nz[r];          // dot(-light_dir, n)
        const float rx = dx[r] - 2.0f * nx[r] * dn, ry = dy[r] - 2.0f * ny[r] * dn, rz = dz[r] - 2.0f * nz[r] * dn;
        const float rdv = fmaxf(0.0f, rx * vx[r] + ry * vy[r] + rz * vz[r]);
        const float spec = 0.5f * (mx[r].x == 0.0f ? 1.0f : __powf(rdv, mx[r].x));
        sr[r] += g_frame.light_col[l][0] * spec + m[r].x * kd;
        sg[r] += g_frame.light_col[l][1] * spec + m[r].y * kd;
        sb[r] += g_frame.light_col[l][2] * spec + m[r].z * kd;
      }
    }
  }
  // continuation: src/main.cpp:43-55 unrolled front to back
#pragma unroll
  for (int r = 0; r < 2; r++) {
    bool push = false;
    RayRec rec;
    if (hit[r]) {
      if (a.r.shadow_mask) a.r.shadow_mask[(size_t)pix[r] * a.r.max_depth + level] = smask[r];
      if (mx[r].y > 0.5f) {                       // reflectivity > 0, decided in double on the host
        const float refl = m[r].w, k = wt[r] * (1.0f - refl);
        cr[r] += k * sr[r]; cg[r] += k * sg[r]; cb[r] += k * sb[r];
        wt[r] *= refl;
        if (level + 1 < a.r.max_depth) {
          d3 o2, d2;
          rtx::reflect_ray(d64v[r], p[r], n[r], 0.001, o2, d2);
          rec.ox = o2.x; rec.oy = o2.y; rec.oz = o2.z; rec.dx = d2.x; rec.dy = d2.y; rec.dz = d2.z;
          rec.pix = pix[r]; rec.wt = wt[r]; rec.ar = cr[r]; rec.ag = cg[r]; rec.ab = cb[r]; rec.pad = 0;
          push = true;
        } else {
          final_[r] = true;                       // depth exhausted: the child contributes black
        }
      } else {
        cr[r] += wt[r] * sr[r]; cg[r] += wt[r] * sg[r]; cb[r] += wt[r] * sb[r];
        final_[r] = true;
      }
    }
    queue_push(push, rec, a.q_out, a.q_out_count);
  }
}

__device__ __forceinline__ void flush_counters(const FastArgs &a, Counters &c, unsigned n_fp64, int level) {
  if (!a.r.counters) return;
  c.fp64 += n_fp64;
  unsigned long long v[6] = {c.closest, c.hits, c.shadow, c.occluded, c.fp64, c.violations};
#pragma unroll
  for (int k = 0; k < 6; k++) v[k] = wsum(v[k]);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&a.r.counters[RT_CNT_CLOSEST], v[0]);
    atomicAdd(&a.r.counters[RT_CNT_HITS], v[1]);
    atomicAdd(&a.r.counters[RT_CNT_SHADOW], v[2]);
    atomicAdd(&a.r.counters[RT_CNT_OCCLUDED], v[3]);
    atomicAdd(&a.r.counters[RT_CNT_FP64], v[4]);
    atomicAdd(&a.r.counters[RT_CNT_VIOLATIONS], v[5]);
    atomicAdd(&a.r.counters[RT_CNT_TESTS], (v[0] + v[2]) * (unsigned long long)a.N);
    if (level < 32) atomicAdd(&a.r.counters[RT_CNT_ALIVE0 + level], v[0]);
  }
}

// Stages the sphere tables of this kernel into shared memory with ONE TMA bulk copy.
__device__ __forceinline__ const float4 *stage_tables(const FastArgs &a, unsigned char *smem, const float4 *gsrc) {
  if (!a.tables_in_smem) return gsrc;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem);
  float4 *dst = reinterpret_cast<float4 *>(smem + kSmemHeader);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, a.table_bytes);
    tma_bulk_g2s(dst, gsrc, a.table_bytes, bar);
  }
  __syncthreads();
  mbar_wait(bar, 0);
  return dst;
}

// ---------------------------------------------------------------------------------------------
// LEVEL 0: camera rays.  Persistent CTAs pull 32x16-pixel tiles from an atomic counter.
// Thread layout inside a tile: warp w covers an 8x8 block (wx = w&3, wy = w>>2), lane = (lx, ly)
// = (lane&7, lane>>3) owns the two vertically adjacent pixels (x, 2*ly) and (x, 2*ly+1).
template <bool kSmem>
__global__ void __launch_bounds__(kThreads, 2) k_primary(const FastArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  int *s_tile = reinterpret_cast<int *>(smem + 16);
  unsigned char *s_rgb = smem + 64;                                  // kTileH x kTileW x 3 = 1536 bytes
  const float4 *tabs = stage_tables(a, smem, a.otab);
  const float4 *cam = tabs, *ltabs = tabs + (size_t)a.npairs * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = a.r.W, rows = a.r.bands.local_rows, depth = a.r.max_depth;
  Counters cnt = {0, 0, 0, 0, 0, 0};
  unsigned n_fp64 = 0;
  for (;;) {
    if (threadIdx.x == 0) *s_tile = (int)atomicAdd(a.tile_counter, 1u);
    __syncthreads();
    const int tile = *s_tile;
    if (tile >= a.ntiles) break;
    const int tx0 = (tile % a.tiles_x) * kTileW, ty0 = (tile / a.tiles_x) * kTileH;
    const int x = tx0 + (warp & 3) * 8 + (lane & 7);
    int lr[2], j[2];
    unsigned pix[2];
    bool live[2];
    float dx[2], dy[2], dz[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      lr[r] = ty0 + (warp >> 2) * 8 + (lane >> 3) * 2 + r;
      live[r] = x < W && lr[r] < rows && depth > 0;
      j[r] = 0; pix[r] = 0; dx[r] = dy[r] = dz[r] = 0.f;
      if (x < W && lr[r] < rows) {
        j[r] = rt_local_to_global_row(a.r.bands, lr[r]);
        pix[r] = (unsigned)lr[r] * (unsigned)W + (unsigned)x;
      }
      if (live[r]) {
        // include/camera.h:21-22 in FP64 (un-normalised), then an FP32 unit vector for the filter
        const d3 v = rtx::add(rtx::add(ldc3(g_frame.fwd), rtx::scale(ldc3(g_frame.right), a.r.su[x])),
                              rtx::scale(ldc3(g_frame.up), a.r.sv[j[r]]));
        const float fx = (float)v.x, fy = (float)v.y, fz = (float)v.z;
        const float inv = rsqrtf(fmaf(fz, fz, fmaf(fy, fy, fx * fx)));
        dx[r] = fx * inv; dy[r] = fy * inv; dz[r] = fz * inv;
        if (a.r.hit_idx) for (int k = 0; k < depth; k++) a.r.hit_idx[(size_t)pix[r] * depth + k] = -2;
        if (a.r.shadow_mask) for (int k = 0; k < depth; k++) a.r.shadow_mask[(size_t)pix[r] * depth + k] = 0u;
      }
    }
    Best best[2];
    best_init(best[0]); best_init(best[1]);
    const double *su = a.r.su, *sv = a.r.sv;
    const int j0 = j[0], j1 = j[1];
    auto ray0 = [&]() { return exact_primary_ray(su, sv, x, j0); };
    auto ray1 = [&]() { return exact_primary_ray(su, sv, x, j1); };
    closest_shared<kSmem>(cam, a.npairs, a.N, dx, dy, dz, live, a.d64, a.r.sph64, ray0, ray1, best, n_fp64);

    bool hit[2], final_[2];
    int idx[2];
    d3 o64[2], d64v[2];
    double t64[2];
    float wt[2] = {1.f, 1.f}, cr[2] = {0.f, 0.f}, cg[2] = {0.f, 0.f}, cb[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; r++) {
      hit[r] = false; final_[r] = false; idx[r] = -1; t64[r] = 0;
      o64[r] = rtx::mk(0, 0, 0); d64v[r] = o64[r];
      if (!live[r]) continue;
      cnt.closest++;
      if (best[r].idx >= 0) {
        ExactRay e = exact_primary_ray(su, sv, x, j[r]);
        double t = best[r].t;
        bool ok = best[r].exact;
        if (!ok) { n_fp64++; ok = exact_sphere(a.r.sph64, best[r].idx, e.o, e.d, e.a, t) && t < 1e20; }
        int bi = best[r].idx;
        if (!ok) { cnt.violations++; bi = exact_bruteforce(a.r.sph64, a.N, e.o, e.d, e.a, t); }
        if (bi >= 0) { hit[r] = true; idx[r] = bi; t64[r] = t; o64[r] = e.o; d64v[r] = e.d; cnt.hits++; }
      }
      if (a.r.hit_idx) a.r.hit_idx[(size_t)pix[r] * depth] = idx[r];
      if (!hit[r]) {                               // sky, src/main.cpp:26-30
        const float ts = 0.5f * (dy[r] + 1.0f);
        cr[r] = (1.0f - ts) + 0.5f * ts; cg[r] = (1.0f - ts) + 0.7f * ts; cb[r] = (1.0f - ts) + ts;
        final_[r] = true;
      }
    }
    shade_and_continue<kSmem>(a, ltabs, hit, idx, o64, d64v, t64, pix, wt, cr, cg, cb, final_, 0, cnt, n_fp64);

    // ---- 8-bit quantise (src/main.cpp:84-86) into the tile staging buffer, then 128-bit row stores
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int ty = (warp >> 2) * 8 + (lane >> 3) * 2 + r, txx = (warp & 3) * 8 + (lane & 7);
      unsigned char *q = s_rgb + (ty * kTileW + txx) * 3;
      const bool blackout = x < W && lr[r] < rows && depth <= 0;
      const bool fin = final_[r] || blackout;
      q[0] = (unsigned char)quant8(fin ? cr[r] : 0.f); q[1] = (unsigned char)quant8(fin ? cg[r] : 0.f);
      q[2] = (unsigned char)quant8(fin ? cb[r] : 0.f);
    }
    __syncthreads();
    {
      const bool full_w = tx0 + kTileW <= W && (W & 15) == 0;
      if (full_w) {
        // 6 x 16-byte stores per 96-byte row segment; pixels still in flight are overwritten by k_bounce
        if (threadIdx.x < kTileH * 6) {
          const int ty = threadIdx.x / 6, seg = threadIdx.x % 6;
          if (ty0 + ty < rows) {
            const uint4 v = *reinterpret_cast<const uint4 *>(s_rgb + ty * kTileW * 3 + seg * 16);
            *reinterpret_cast<uint4 *>(a.r.rgb + ((size_t)(ty0 + ty) * W + tx0) * 3 + seg * 16) = v;
          }
        }
      } else {
        for (int k = threadIdx.x; k < kTileH * kTileW; k += kThreads) {
          const int ty = k / kTileW, txx = k % kTileW;
          if (tx0 + txx < W && ty0 + ty < rows) {
            unsigned char *o = a.r.rgb + ((size_t)(ty0 + ty) * W + tx0 + txx) * 3;
            const unsigned char *q = s_rgb + k * 3;
            o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
          }
        }
      }
    }
    __syncthreads();
  }
  flush_counters(a, cnt, n_fp64, 0);
}

// ---------------------------------------------------------------------------------------------
// LEVEL >= 1: reflected rays from the queue, two per thread, 512 per CTA chunk.
template <bool kSmem>
__global__ void __launch_bounds__(kThreads, 2) k_bounce(const FastArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  int *s_chunk = reinterpret_cast<int *>(smem + 16);
  if (*a.q_in_count == 0u) return;                  // nothing survived to this level
  // tables: [L light tables][general table], contiguous in global memory in that order
  const float4 *tabs = stage_tables(a, smem, a.otab + (size_t)a.npairs * 2);
  const float4 *ltabs = tabs, *gen = tabs + (size_t)a.L * a.npairs * 2;
  const unsigned nq = *a.q_in_count;
  const int nchunks = (int)((nq + 2 * kThreads - 1) / (2 * kThreads));
  const int depth = a.r.max_depth, level = a.level;
  Counters cnt = {0, 0, 0, 0, 0, 0};
  unsigned n_fp64 = 0;
  for (;;) {
    if (threadIdx.x == 0) *s_chunk = (int)atomicAdd(a.chunk_counter, 1u);
    __syncthreads();
    const int chunk = *s_chunk;
    __syncthreads();
    if (chunk >= nchunks) break;
    bool live[2];
    unsigned qi[2], pix[2];
    float ox[2], oy[2], oz[2], dx[2], dy[2], dz[2];
    float wt[2], cr[2], cg[2], cb[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      // interleave so that the two rays of a thread are neighbours in the queue (coherent)
      qi[r] = (unsigned)chunk * 2u * kThreads + 2u * threadIdx.x + r;
      live[r] = qi[r] < nq;
      ox[r] = oy[r] = oz[r] = dx[r] = dy[r] = dz[r] = 0.f; wt[r] = cr[r] = cg[r] = cb[r] = 0.f; pix[r] = 0;
      if (live[r]) {
        const RayRec &q = a.q_in[qi[r]];
        // recentred FP32 origin and FP32 direction for the filter (exact values stay in the record)
        ox[r] = (float)(q.ox - a.c0[0]); oy[r] = (float)(q.oy - a.c0[1]); oz[r] = (float)(q.oz - a.c0[2]);
        dx[r] = (float)q.dx; dy[r] = (float)q.dy; dz[r] = (float)q.dz;
        pix[r] = q.pix; wt[r] = q.wt; cr[r] = q.ar; cg[r] = q.ag; cb[r] = q.ab;
      }
    }
    Best best[2];
    best_init(best[0]); best_init(best[1]);
    const RayRec *qin = a.q_in;
    const unsigned q0 = qi[0], q1 = qi[1];
    auto mkray = [&](unsigned k) {
      ExactRay e;
      const RayRec &q = qin[k];
      e.o = rtx::mk(q.ox, q.oy, q.oz); e.d = rtx::mk(q.dx, q.dy, q.dz); e.a = rtx::dot(e.d, e.d);
      return e;
    };
    auto ray0 = [&]() { return mkray(q0); };
    auto ray1 = [&]() { return mkray(q1); };
    closest_general<kSmem>(gen, a.npairs, a.N, ox, oy, oz, dx, dy, dz, live, a.d64, a.gS2, a.r.sph64, ray0, ray1, best, n_fp64);

    bool hit[2], final_[2];
    int idx[2];
    d3 o64[2], d64v[2];
    double t64[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      hit[r] = false; final_[r] = false; idx[r] = -1; t64[r] = 0;
      o64[r] = rtx::mk(0, 0, 0); d64v[r] = o64[r];
      if (!live[r]) continue;
      cnt.closest++;
      if (best[r].idx >= 0) {
        ExactRay e = mkray(qi[r]);
        double t = best[r].t;
        bool ok = best[r].exact;
        if (!ok) { n_fp64++; ok = exact_sphere(a.r.sph64, best[r].idx, e.o, e.d, e.a, t) && t < 1e20; }
        int bi = best[r].idx;
        if (!ok) { cnt.violations++; bi = exact_bruteforce(a.r.sph64, a.N, e.o, e.d, e.a, t); }
        if (bi >= 0) { hit[r] = true; idx[r] = bi; t64[r] = t; o64[r] = e.o; d64v[r] = e.d; cnt.hits++; }
      }
      if (a.r.hit_idx) a.r.hit_idx[(size_t)pix[r] * depth + level] = idx[r];
      if (!hit[r]) {
        const float ts = 0.5f * (dy[r] + 1.0f);
        cr[r] += wt[r] * ((1.0f - ts) + 0.5f * ts); cg[r] += wt[r] * ((1.0f - ts) + 0.7f * ts); cb[r] += wt[r] * ((1.0f - ts) + ts);
        final_[r] = true;
      }
    }
    shade_and_continue<kSmem>(a, ltabs, hit, idx, o64, d64v, t64, pix, wt, cr, cg, cb, final_, level, cnt, n_fp64);
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (live[r] && final_[r]) {
        unsigned char *o = a.r.rgb + (size_t)pix[r] * 3;
        o[0] = (unsigned char)quant8(cr[r]); o[1] = (unsigned char)quant8(cg[r]); o[2] = (unsigned char)quant8(cb[r]);
      }
    }
  }
  flush_counters(a, cnt, n_fp64, level);
}

}  // namespace rtf
#endif
