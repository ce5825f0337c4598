// kernels_fast.cuh -- "mode 0": the production path.  sm_100a only.
//
// Work decomposition (wavefront with compacted queues, warps schedule themselves):
//   k_primary   persistent CTAs; every WARP pulls 16x4-pixel tiles from an atomic counter (two
//               vertically adjacent pixels per lane).  Camera rays -> closest hit (FP32 filter over
//               the CAMERA table, FP64 decide) -> Phong + one any-hit shadow query per light (FP32
//               filter over that LIGHT's table) -> final 8-bit pixel (staged per warp, 128-bit
//               stores), or a reflected ray appended to the ray queue with warp-ballot compaction.
//   k_bounce    level 1: consumes the queue, 64 rays per warp fetch, general-origin closest hit,
//               same shading, appends survivors to the other queue.
//               tail (levels >= 2, one launch): each lane follows its ray to termination, the
//               queue record is updated in place instead of being re-queued.
// Sphere tables are staged once per CTA into shared memory with ONE TMA bulk copy
// (cp.async.bulk + mbarrier) when they fit, otherwise they are read through L1/L2.
//
// Shared-origin tables (camera, each light) are SORTED by the distance of the sphere's nearest
// point from that origin; a query stops at the first 8-sphere group that lies entirely beyond
// its cutoff (current best hit / distance to the shaded point), so "any hit" really is early out.
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
constexpr int kWarps = kThreads / 32;
constexpr int kWTileW = 16, kWTileH = 4;  // pixels per warp tile (32 lanes x 2 pixels)
constexpr int kGroupPairs = 4;            // sphere pairs per fast-path group (8 spheres)
constexpr int kSmemHeader = 2048;         // [0,8) mbarrier, [64, 64+8*192) per-warp RGB staging
constexpr float kEps = 0.001f;            // EPSILON, include/ray_math_constants.h:22
constexpr unsigned kSign = 0x80000000u;
constexpr unsigned kFull = 0xffffffffu;

struct __align__(16) RayRec {             // one queued reflected ray (80 bytes)
  double ox, oy, oz, dx, dy, dz;          // exact FP64 origin / unit direction (src/main.cpp:45-48)
  unsigned pix;                           // local pixel index lr*W + x
  float wt;                               // product of reflectivities so far
  float ar, ag, ab;                       // colour accumulated so far (front to back)
  unsigned pad;
};

struct FastArgs {
  RtRenderArgs r;
  const unsigned char *tabs;  // (1+L) shared-origin tables (tstride bytes each: pairs | gmin | perm), then the general table
  int npairs, ngroups;        // npairs is a multiple of kGroupPairs; ngroups = npairs / kGroupPairs
  int N, L;
  unsigned tstride, gmin_off, perm_off, inv_off;
  unsigned stage_bytes;       // bytes this kernel stages into shared memory
  float d64;                  // absolute slack covering FP64 rounding / geometry (delta64)
  float gS2;                  // squared radius bound S^2 of the recentred scene (general filter)
  float g_dtmax;              // 16u*S: bound of |fl32 dot - true| for the general filter
  double c0[3];               // recentring offset of the general table
  int wtiles_x, nwtiles;
  unsigned int *tile_counter;
  RayRec *q_out; unsigned int *q_out_count;
  RayRec *q_in; const unsigned int *q_in_count;
  unsigned int *chunk_counter;
  int level;
  int tables_in_smem;
};

// One shared-origin table: sphere pairs in sorted order, per-group minimum distance, original indices
struct Tab {
  const float4 *pairs;   // 2 float4 per pair: (x0,x1,y0,y1) (z0,z1,w0,w1)
  const float *gmin;     // per group of 8 spheres: lower bound of |oc| - r over the group (ascending)
  const int *perm;       // original sphere index per sorted slot, -1 = padding
  const int *inv;        // sorted slot of each original sphere index; bit 30: the origin is strictly outside it
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
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float wmaxf(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ unsigned quant8(float c) { return (unsigned)(int)(255.99f * fminf(1.0f, c)); }
__device__ __forceinline__ int warp_fetch(unsigned int *counter) {
  int v = 0;
  if ((threadIdx.x & 31) == 0) v = (int)atomicAdd(counter, 1u);
  return __shfl_sync(kFull, v, 0);
}

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

// Stages `bytes` from gsrc into shared memory behind the header with ONE TMA bulk copy.
__device__ __forceinline__ void stage_tables(unsigned char *smem, const unsigned char *gsrc, unsigned bytes) {
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, bytes);
    tma_bulk_g2s(smem + kSmemHeader, gsrc, bytes, bar);
  }
  __syncthreads();
  mbar_wait(bar, 0);
}

__device__ __forceinline__ Tab tab_at(const unsigned char *base, const FastArgs &a, int t) {
  const unsigned char *p = base + (size_t)t * a.tstride;
  Tab T;
  T.pairs = reinterpret_cast<const float4 *>(p);
  T.gmin = reinterpret_cast<const float *>(p + a.gmin_off);
  T.perm = reinterpret_cast<const int *>(p + a.perm_off);
  T.inv = reinterpret_cast<const int *>(p + a.inv_off);
  return T;
}

// ---------------------------------------------------------------------------------------------
// exact (FP64) deciders.  Everything here is OUT OF LINE on purpose: it is rare, register hungry
// and large (IEEE double sqrt / div expand to dozens of instructions); one copy per kernel keeps
// the instruction footprint of the kernels inside the instruction cache.
struct ExactRay { d3 o, d; double a; };

// Where a query's exact FP64 ray comes from: the camera (pixel x, row j) or a queue record.
struct RaySrc { const double *su, *sv; int x, j; const RayRec *rec; };

__device__ __noinline__ ExactRay exact_ray(RaySrc s) {
  ExactRay e;
  if (s.rec) {
    e.o = rtx::mk(s.rec->ox, s.rec->oy, s.rec->oz);
    e.d = rtx::mk(s.rec->dx, s.rec->dy, s.rec->dz);
  } else {
    e.o = ldc3(g_frame.cam_pos);
    e.d = rtx::camera_dir(ldc3(g_frame.fwd), ldc3(g_frame.right), ldc3(g_frame.up), s.su[s.x], s.sv[s.j]);
  }
  e.a = rtx::dot(e.d, e.d);
  return e;
}
__device__ __noinline__ bool exact_sphere(const double4 *sph64, int idx, d3 o, d3 d, double a, double &t) {
  double4 s = ld_sph64(&sph64[idx]);
  return rtx::intersect(o, d, a, rtx::mk(s.x, s.y, s.z), s.w, t);
}
// include/scene.h:70-85 for ONE sphere: the reference's own formula on the reference's own shadow ray
__device__ __noinline__ bool exact_shadow_sphere(const double4 *sph64, int idx, d3 p, int light) {
  d3 o, d; double ldist, tt;
  rtx::shadow_ray(p, ldc3(g_frame.light_pos[light]), 0.001, o, d, ldist);
  return exact_sphere(sph64, idx, o, d, rtx::dot(d, d), tt) && tt < 1e20 && tt < ldist;
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
// src/main.cpp:32,35 : hit point and unit normal
struct HitGeom { d3 p, n; };
__device__ __noinline__ HitGeom hit_geometry(const double4 *sph64, int idx, d3 o, d3 d, double t) {
  const double4 s = ld_sph64(&sph64[idx]);
  HitGeom h;
  h.p = rtx::hit_point(o, d, t);
  h.n = rtx::normal_at(h.p, rtx::mk(s.x, s.y, s.z));
  return h;
}
// src/main.cpp:45-48 : the reflected ray as the Ray ctor stores it
__device__ __noinline__ void reflected_ray(d3 d, d3 p, d3 n, RayRec *rec) {
  d3 o2, d2;
  rtx::reflect_ray(d, p, n, 0.001, o2, d2);
  rec->ox = o2.x; rec->oy = o2.y; rec->oz = o2.z; rec->dx = d2.x; rec->dy = d2.y; rec->dz = d2.z;
}

// ---------------------------------------------------------------------------------------------
// closest-hit bookkeeping: best candidate as an FP32 bracket, exact FP64 t only when needed.
// All comparisons are order independent: (t, index) lexicographic, i.e. the reference's strict '<'
// scan in ascending index order (include/scene.h:47-56), whatever order spheres are visited in.
struct Best {
  float lo, hi;
  int idx;
  int nfp64;      // FP64 sphere evaluations spent on this query
  double t;       // valid iff exact
  bool exact;
};
__device__ __forceinline__ void best_init(Best &b) { b.lo = 3.0e38f; b.hi = 3.0e38f; b.idx = -1; b.t = 1e20; b.exact = false; b.nfp64 = 0; }
__device__ __forceinline__ void best_set_exact(Best &b, int idx, double t) {
  b.idx = idx; b.t = t; b.exact = true; b.lo = f_rd(t); b.hi = f_ru(t);
}

__device__ __noinline__ Best closest_consider(Best b, int i, int status, float lo, float hi, const double4 *sph64, RaySrc src) {
  if (status == RT_MISS) return b;
  if (status == RT_HIT) {
    if (b.idx < 0) { b.lo = lo; b.hi = hi; b.idx = i; b.exact = false; return b; }
    if (lo > b.hi) return b;                      // strictly farther
    if (hi < b.lo) { b.lo = lo; b.hi = hi; b.idx = i; b.exact = false; return b; }
  }
  // brackets touch, or the sphere itself is ambiguous: decide in FP64, lowest index wins ties
  const ExactRay e = exact_ray(src);
  double tn;
  b.nfp64++;
  const bool hn = exact_sphere(sph64, i, e.o, e.d, e.a, tn);
  if (!hn || !(tn < 1e20)) return b;              // INFINITY_DOUBLE init of include/scene.h:42
  if (b.idx >= 0 && !b.exact) {
    const float tl = f_rd(tn), th = f_ru(tn);
    if (tl > b.hi) return b;
    if (th < b.lo) { best_set_exact(b, i, tn); return b; }
    double tb;
    b.nfp64++;
    const bool hb = exact_sphere(sph64, b.idx, e.o, e.d, e.a, tb);
    if (hb) best_set_exact(b, b.idx, tb); else { const int k = b.nfp64; best_init(b); b.nfp64 = k; }   // (else: filter violation, caught later)
  }
  if (b.idx < 0 || tn < b.t || (tn == b.t && i < b.idx)) best_set_exact(b, i, tn);
  return b;
}

// scalar re-evaluation of one sphere of a shared-origin table (same operations as the fast path)
__device__ __forceinline__ void shared_origin_eval(float ocx, float ocy, float ocz, float ncc, float dx, float dy, float dz,
                                                   float &tca, float &Dp) {
  tca = __fmul_rn(ocx, dx);
  tca = __fmaf_rn(ocy, dy, tca);
  tca = __fmaf_rn(ocz, dz, tca);
  Dp = __fmaf_rn(tca, tca, ncc);
}

// Brackets the roots of sphere (oc, ncc) of a shared-origin table.  False = discriminant sign uncertain.
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

// ---- slow paths: one flagged sphere PAIR for ONE ray, out of line --------------------------------
__device__ __noinline__ Best slow_closest_shared(Best b, const float4 *pairs, const int *perm, int pi, float dx, float dy, float dz,
                                                 float d64, const double4 *sph64, RaySrc src) {
  const float4 A = pairs[2 * pi], B = pairs[2 * pi + 1];
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    const int i = perm[2 * pi + h];
    if (i < 0) continue;
    const float ocx = h ? A.y : A.x, ocy = h ? A.w : A.z, ocz = h ? B.y : B.x, ncc = h ? B.w : B.z;
    float tca, Dp;
    shared_origin_eval(ocx, ocy, ocz, ncc, dx, dy, dz, tca, Dp);
    if (!(Dp >= 0.0f)) continue;
    Roots rt;
    int status = RT_AMBIG;
    float lo = 0, hi = 0;
    if (shared_origin_roots(ocx, ocy, ocz, ncc, tca, Dp, d64, rt)) status = select_root(rt, lo, hi);
    b = closest_consider(b, i, status, lo, hi, sph64, src);
  }
  return b;
}

// returns bit 0 = an occluder was found in this pair, bits 1.. = FP64 evaluations spent
__device__ __noinline__ int slow_shadow(const float4 *pairs, const int *perm, int pi, float dx, float dy, float dz, float so, float m,
                                        int self, float cosl, const double *p3, int light, float d64, const double4 *sph64) {
  const float so_lo = so - m, so_hi = so + m, e_lo = -kEps - m, e_hi = -kEps + m;
  const float4 A = pairs[2 * pi], B = pairs[2 * pi + 1];
  int n64 = 0, found = 0;
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    const int i = perm[2 * pi + h];
    if (i < 0) continue;
    if (i == self && cosl > 1e-3f) continue;
    const float ocx = h ? A.y : A.x, ocy = h ? A.w : A.z, ocz = h ? B.y : B.x, ncc = h ? B.w : B.z;
    float tca, Dp;
    shared_origin_eval(ocx, ocy, ocz, ncc, dx, dy, dz, tca, Dp);
    if (!(Dp >= 0.0f)) continue;
    Roots rt;
    if (shared_origin_roots(ocx, ocy, ocz, ncc, tca, Dp, d64, rt)) {
      const bool no = (rt.n_lo > so_hi) || (rt.f_hi < e_lo) || (rt.n_hi < e_lo && rt.f_lo > so_hi);
      const bool yes = (rt.f_lo > e_hi && rt.f_hi < so_lo) || (rt.f_lo > so_hi && rt.n_lo > e_hi && rt.n_hi < so_lo);
      if (no) continue;
      if (yes) { found = 1; break; }
    }
    n64++;
    if (exact_shadow_sphere(sph64, i, rtx::mk(p3[0], p3[1], p3[2]), light)) { found = 1; break; }
  }
  return found | (n64 << 1);
}

__device__ __noinline__ Best slow_closest_general(Best b, const float4 *pairs, int pi, int N, float ox, float oy, float oz, float dx,
                                                  float dy, float dz, float d64, float gS2, const double4 *sph64, RaySrc src) {
  const float sS = __fsqrt_ru(gS2);
  const float4 A = pairs[2 * pi], B = pairs[2 * pi + 1];
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    const int i = 2 * pi + h;
    if (i >= N) break;
    const float cx = h ? A.y : A.x, cy = h ? A.w : A.z, cz = h ? B.y : B.x, rho = h ? B.w : B.z;
    // explicit-margin evaluation (independent of the inflation tricks of the fast path)
    const float x = __fsub_rn(cx, ox), y = __fsub_rn(cy, oy), z = __fsub_rn(cz, oz);
    float tca = __fmul_rn(x, dx); tca = __fmaf_rn(y, dy, tca); tca = __fmaf_rn(z, dz, tca);
    const float oc2 = __fmaf_ru(z, z, __fmaf_ru(y, y, __fmul_ru(x, x)));
    // |X - oc*| <= u(2S + |oc|)  =>  D error <= u (8.1 S |oc| + 16 |oc|^2 + 4 rho) + d64   
    const float ocn = __fsqrt_ru(oc2);
    const float Eg = __fadd_ru(__fmul_ru(5.9604645e-8f, __fmaf_ru(8.2f * sS, ocn, __fmaf_ru(16.5f, oc2, 4.5f * fabsf(rho)))), d64);
    const float Dc = __fmaf_rn(tca, tca, __fsub_rn(rho, oc2));      // rho = r^2 + margins
    const float Dhi = __fadd_ru(Dc, Eg);
    if (!(Dhi >= 0.0f)) continue;
    Roots rt;
    int status = RT_AMBIG;
    float lo = 0, hi = 0;
    const float dt = __fmul_ru(RT_ETA * 1.001f, __fadd_ru(ocn, sS));
    // rho' - r^2 = 40u r^2 + 12u S r + 64u^2 S^2 + d64 (host), bounded here from rho' itself
    const float rm = __fadd_ru(__fmul_ru(5.9604645e-8f, __fmaf_ru(12.5f * sS, __fsqrt_ru(fabsf(rho)), __fmaf_ru(41.0f, fabsf(rho), 1e-4f * gS2))), d64);
    if (bracket_roots(tca, Dhi, __fadd_ru(__fmul_ru(2.0f, Eg), rm), dt, rt)) status = select_root(rt, lo, hi);
    b = closest_consider(b, i, status, lo, hi, sph64, src);
  }
  return b;
}

// The packed FP32 test of up to 32 sphere PAIRS (one "chunk") of a shared-origin table against the
// two rays of a lane.  For every pair and ray one bit is shifted into a history word: 1 = neither
// sphere of the pair can be hit (D' = (oc.d)^2 + ncc < 0 for both), 0 = flagged.  No branch, no vote:
// flagged pairs are resolved afterwards, all lanes together (see the drain loops below).
constexpr int kChunkPairs = 32;
__device__ __forceinline__ unsigned push_sign(unsigned hist, unsigned v) { return __funnelshift_l(v, hist, 1); }

__device__ __forceinline__ void chunk_test_shared(const float4 *__restrict__ pairs, int p0, int np, const float2 (&dx)[2],
                                                  const float2 (&dy)[2], const float2 (&dz)[2], unsigned &h0, unsigned &h1) {
  h0 = kFull; h1 = kFull;
#pragma unroll 1
  for (int p = p0; p < p0 + np; p += kGroupPairs) {
#pragma unroll
    for (int k = 0; k < kGroupPairs; k++) {
      const float4 A = pairs[2 * (p + k)], B = pairs[2 * (p + k) + 1];
      const float2 X = make_float2(A.x, A.y), Y = make_float2(A.z, A.w), Z = make_float2(B.x, B.y), Wv = make_float2(B.z, B.w);
      float2 t0 = __fmul2_rn(X, dx[0]); t0 = __ffma2_rn(Y, dy[0], t0); t0 = __ffma2_rn(Z, dz[0], t0);
      float2 t1 = __fmul2_rn(X, dx[1]); t1 = __ffma2_rn(Y, dy[1], t1); t1 = __ffma2_rn(Z, dz[1], t1);
      const float2 D0 = __ffma2_rn(t0, t0, Wv), D1 = __ffma2_rn(t1, t1, Wv);
      h0 = push_sign(h0, fbits(D0.x) & fbits(D0.y));
      h1 = push_sign(h1, fbits(D1.x) & fbits(D1.y));
    }
  }
}
// history word -> flagged-pair bits (bit i <-> pair p0 + np - 1 - i)
__device__ __forceinline__ unsigned flagged_bits(unsigned hist, int np, bool live) {
  const unsigned pm = np >= 32 ? kFull : ((1u << np) - 1u);
  return live ? (~hist & pm) : 0u;
}

// ---------------------------------------------------------------------------------------------
// CLOSEST HIT, shared origin (camera table, sorted by nearest-point distance).  Two rays per lane.
// The query is resumable over pair ranges so that a table larger than shared memory can be streamed
// through it tile by tile: `pairs` is indexed with ABSOLUTE pair numbers (a tile buffer is passed as
// buffer - 2*first_pair), gmin / perm may live in global memory.
struct ClosestQ {
  Best best[2];
  float wcut;                                        // warp-uniform: farthest cutoff of any live ray
};
__device__ __forceinline__ void closest_begin(ClosestQ &q) { best_init(q.best[0]); best_init(q.best[1]); q.wcut = 3.0e38f; }

// returns false once every remaining sphere of the (sorted) table is beyond every ray's best hit
__device__ __forceinline__ bool closest_shared_range(ClosestQ &q, const float4 *__restrict__ pairs, const float *gmin, const int *perm,
                                                     int pbeg, int pend, const float (&dx)[2], const float (&dy)[2],
                                                     const float (&dz)[2], const bool (&live)[2], float d64, const double4 *sph64,
                                                     const RaySrc (&src)[2]) {
  const float2 dx2[2] = {make_float2(dx[0], dx[0]), make_float2(dx[1], dx[1])};
  const float2 dy2[2] = {make_float2(dy[0], dy[0]), make_float2(dy[1], dy[1])};
  const float2 dz2[2] = {make_float2(dz[0], dz[0]), make_float2(dz[1], dz[1])};
#pragma unroll 1
  for (int p0 = pbeg; p0 < pend; p0 += kChunkPairs) {
    if (gmin[p0 / kGroupPairs] > q.wcut) return false;
    const int np = min(kChunkPairs, pend - p0);
    unsigned h0, h1;
    chunk_test_shared(pairs, p0, np, dx2, dy2, dz2, h0, h1);
    const unsigned f[2] = {flagged_bits(h0, np, live[0]), flagged_bits(h1, np, live[1])};
    if (__any_sync(kFull, (f[0] | f[1]) != 0u)) {
      // drain: every lane resolves its own flagged pairs, nearest first, one per iteration
#pragma unroll
      for (int r = 0; r < 2; r++) {
        unsigned fr = f[r];
        while (__any_sync(kFull, fr != 0u)) {
          if (fr != 0u) {
            const int bit = 31 - __clz(fr);
            const int pi = p0 + np - 1 - bit;
            fr &= ~(1u << bit);
            if (gmin[pi / kGroupPairs] > q.best[r].hi) fr = 0u;       // sorted: the rest is farther still
            else q.best[r] = slow_closest_shared(q.best[r], pairs, perm, pi, dx[r], dy[r], dz[r], d64, sph64, src[r]);
          }
        }
      }
      q.wcut = wmaxf(fmaxf(live[0] ? q.best[0].hi : -3.0e38f, live[1] ? q.best[1].hi : -3.0e38f));
    }
  }
  return true;
}

__device__ __forceinline__ void closest_shared(const Tab T, int npairs, const float (&dx)[2], const float (&dy)[2],
                                               const float (&dz)[2], const bool (&live)[2], float d64, const double4 *sph64,
                                               const RaySrc (&src)[2], Best (&best)[2]) {
  ClosestQ q;
  closest_begin(q);
  closest_shared_range(q, T.pairs, T.gmin, T.perm, 0, npairs, dx, dy, dz, live, d64, sph64, src);
  best[0] = q.best[0]; best[1] = q.best[1];
}

// ---------------------------------------------------------------------------------------------
// CLOSEST HIT, general origin (bounce rays).  Table pair = (cx0,cx1,cy0,cy1) (cz0,cz1,rho0,rho1)
// with recentred centres, index order.  Per sphere: X = c - o, tu = X.d(1+24u) + dtmax (an UPPER
// bound of the true centre projection), q = |X|^2 - rho' (>= 0 => origin strictly outside),
// D' = tu^2 - q.  A sphere is skipped when D' < 0, or when it lies behind an origin that is outside
// it (tu < 0 and q >= 0): that removes the sphere the ray just left without any extra arithmetic.
__device__ __forceinline__ void closest_general(const float4 *__restrict__ pairs, int npairs, int N, const float (&ox)[2],
                                                const float (&oy)[2], const float (&oz)[2], const float (&dx)[2],
                                                const float (&dy)[2], const float (&dz)[2], const bool (&live)[2], float d64,
                                                float gS2, float dtmax, const double4 *sph64, const RaySrc (&src)[2],
                                                Best (&best)[2]) {
  const float kInfl = 1.0f + 24.0f * 5.9604645e-8f;
  float2 nox[2], noy[2], noz[2], idx2[2], idy2[2], idz2[2];
  const float2 dtm = make_float2(dtmax, dtmax);
#pragma unroll
  for (int r = 0; r < 2; r++) {
    nox[r] = make_float2(-ox[r], -ox[r]); noy[r] = make_float2(-oy[r], -oy[r]); noz[r] = make_float2(-oz[r], -oz[r]);
    float ix = __fmul_rn(dx[r], kInfl), iy = __fmul_rn(dy[r], kInfl), iz = __fmul_rn(dz[r], kInfl);
    idx2[r] = make_float2(ix, ix); idy2[r] = make_float2(iy, iy); idz2[r] = make_float2(iz, iz);
  }
#pragma unroll 1
  for (int p0 = 0; p0 < npairs; p0 += kChunkPairs) {
    const int np = min(kChunkPairs, npairs - p0);
    unsigned h[2] = {kFull, kFull};
#pragma unroll 1
    for (int p = p0; p < p0 + np; p += 2) {
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const float4 A = pairs[2 * (p + k)], B = pairs[2 * (p + k) + 1];
        const float2 CX = make_float2(A.x, A.y), CY = make_float2(A.z, A.w), CZ = make_float2(B.x, B.y);
        const float2 NR = make_float2(-B.z, -B.w);
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const float2 x = __fadd2_rn(CX, nox[r]), y = __fadd2_rn(CY, noy[r]), z = __fadd2_rn(CZ, noz[r]);
          float2 t = __ffma2_rn(x, idx2[r], dtm); t = __ffma2_rn(y, idy2[r], t); t = __ffma2_rn(z, idz2[r], t);
          float2 q = __ffma2_rn(x, x, NR); q = __ffma2_rn(y, y, q); q = __ffma2_rn(z, z, q);
          const float2 D = __ffma2_rn(t, t, make_float2(-q.x, -q.y));
          // rejected  <=>  D' < 0  or  (tu < 0 and q >= 0)
          h[r] = push_sign(h[r], (fbits(D.x) | (fbits(t.x) & ~fbits(q.x))) & (fbits(D.y) | (fbits(t.y) & ~fbits(q.y))));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
      unsigned fr = flagged_bits(h[r], np, live[r]);
      while (__any_sync(kFull, fr != 0u)) {
        if (fr != 0u) {
          const int bit = 31 - __clz(fr);
          const int pi = p0 + np - 1 - bit;
          fr &= ~(1u << bit);
          best[r] = slow_closest_general(best[r], pairs, pi, N, ox[r], oy[r], oz[r], dx[r], dy[r], dz[r], d64, gS2, sph64, src[r]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// SHADOW (any hit, early out), light table sorted by distance from the light.  dl = unit vector FROM
// THE LIGHT TOWARDS the shaded point, so = distance light -> shadow-ray origin (= |L-p| - EPS), both
// FP32 with known error.  Roots s are measured from the light; the reference's t = so - s
// (include/scene.h:70-85):   occluded  <=>  (-EPS < s2 <= so)  or  (s2 > so and -EPS < s1 <= so).
// self[r] / cosl[r]: the sphere the point lies on and n.light_dir there.  Its surface passes through
// the point, so the filter flags it for every query; on its lit side it cannot occlude (the shadow
// origin is outside it and moving away), so its flag is cleared up front unless its pair partner
// is a candidate too.  p64[r] points at the exact hit point (only read if FP64 is needed).
struct ShadowQ {
  bool open[2], occ[2];                              // still undecided / found an occluder
  float m[2], cut[2], wcut;
  int sslot[2];                                      // slot of the lit self sphere in this light's table, or -1
};
__device__ __forceinline__ void shadow_begin(ShadowQ &q, const int *inv, const float (&so)[2], const bool (&want)[2],
                                             const int (&self)[2], const float (&cosl)[2]) {
#pragma unroll
  for (int r = 0; r < 2; r++) {
    q.open[r] = want[r]; q.occ[r] = false;
    q.m[r] = __fmaf_ru(1.9073486e-6f, so[r] + kEps, 1e-7f);   // 2^-19 |L-p|: covers the FP32 length error
    q.cut[r] = want[r] ? so[r] + q.m[r] : -3.0e38f;           // nothing farther from the light can matter
    q.sslot[r] = (want[r] && cosl[r] > 1e-3f) ? (inv[self[r]] & 0x3fffffff) : -1;
  }
  q.wcut = wmaxf(fmaxf(q.cut[0], q.cut[1]));
}

// returns false once the warp needs nothing further from this (sorted) table
__device__ __forceinline__ bool shadow_range(ShadowQ &q, const float4 *__restrict__ pairs, const float *gmin, const int *perm, int pbeg,
                                             int pend, int light, const float (&dx)[2], const float (&dy)[2], const float (&dz)[2],
                                             const float (&so)[2], const int (&self)[2], const float (&cosl)[2],
                                             const double *const (&p64)[2], float d64, const double4 *sph64, int &n_fp64) {
  const float2 dx2[2] = {make_float2(dx[0], dx[0]), make_float2(dx[1], dx[1])};
  const float2 dy2[2] = {make_float2(dy[0], dy[0]), make_float2(dy[1], dy[1])};
  const float2 dz2[2] = {make_float2(dz[0], dz[0]), make_float2(dz[1], dz[1])};
#pragma unroll 1
  for (int p0 = pbeg; p0 < pend; p0 += kChunkPairs) {
    if (gmin[p0 / kGroupPairs] > q.wcut) return false;
    const int np = min(kChunkPairs, pend - p0);
    unsigned h0, h1;
    chunk_test_shared(pairs, p0, np, dx2, dy2, dz2, h0, h1);
    unsigned f[2] = {flagged_bits(h0, np, q.open[0]), flagged_bits(h1, np, q.open[1])};
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int sp = q.sslot[r] >> 1;                // pair of the lit self sphere (or -1)
      if (sp >= p0 && sp < p0 + np) {
        const int ph = (q.sslot[r] & 1) ^ 1;         // partner = the other half of the pair
        const float4 A = pairs[2 * sp], B = pairs[2 * sp + 1];
        float tca, Dp;
        shared_origin_eval(ph ? A.y : A.x, ph ? A.w : A.z, ph ? B.y : B.x, ph ? B.w : B.z, dx[r], dy[r], dz[r], tca, Dp);
        if (!(Dp >= 0.0f)) f[r] &= ~(1u << (p0 + np - 1 - sp));
      }
    }
    if (__any_sync(kFull, (f[0] | f[1]) != 0u)) {
#pragma unroll
      for (int r = 0; r < 2; r++) {
        unsigned fr = f[r];
        while (__any_sync(kFull, fr != 0u)) {
          if (fr != 0u) {
            const int bit = 31 - __clz(fr);
            const int pi = p0 + np - 1 - bit;
            fr &= ~(1u << bit);
            if (gmin[pi / kGroupPairs] > q.cut[r]) {
              fr = 0u;                                // sorted: everything after this pair is farther still
            } else {
              const int rc = slow_shadow(pairs, perm, pi, dx[r], dy[r], dz[r], so[r], q.m[r], self[r], cosl[r], p64[r], light, d64, sph64);
              n_fp64 += rc >> 1;
              if (rc & 1) { q.occ[r] = true; q.open[r] = false; fr = 0u; }
            }
          }
        }
      }
      q.wcut = wmaxf(fmaxf(q.open[0] ? q.cut[0] : -3.0e38f, q.open[1] ? q.cut[1] : -3.0e38f));   // decided rays stop holding the warp
      if (q.wcut < -1.0e38f) return false;
    }
  }
  return true;
}

__device__ __forceinline__ void shadow_light(const Tab T, int npairs, int light, const float (&dx)[2], const float (&dy)[2],
                                             const float (&dz)[2], const float (&so)[2], const bool (&want)[2],
                                             const int (&self)[2], const float (&cosl)[2], const double *const (&p64)[2], float d64,
                                             const double4 *sph64, bool (&occ)[2], int &n_fp64) {
  ShadowQ q;
  shadow_begin(q, T.inv, so, want, self, cosl);
  shadow_range(q, T.pairs, T.gmin, T.perm, 0, npairs, light, dx, dy, dz, so, self, cosl, p64, d64, sph64, n_fp64);
  occ[0] = q.occ[0]; occ[1] = q.occ[1];
}

// ---------------------------------------------------------------------------------------------
// warp-ballot compaction: rays still alive are appended densely to the next level's queue
__device__ __forceinline__ void queue_push(bool want, const RayRec &rec, RayRec *q, unsigned int *count) {
  const unsigned mk = __ballot_sync(kFull, want);
  if (mk == 0) return;
  const int lane = threadIdx.x & 31, leader = __ffs(mk) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(count, (unsigned)__popc(mk));
  base = __shfl_sync(kFull, base, leader);
  if (want) q[base + __popc(mk & ((1u << lane) - 1u))] = rec;
}

struct Counters {
  unsigned long long closest, hits, shadow, occluded, fp64, violations;
};

// ---------------------------------------------------------------------------------------------
// Shading of up to two hits per lane (shared by both kernels).
//   in : hit[r], sphere index, exact FP64 ray (o, d) and t of the hit, carried acc/wt
//   out: final[r] (pixel finished, colour in cr/cg/cb) or cont[r] + rec[r] (reflected ray).
__device__ __forceinline__ void shade_hits(const FastArgs &a, const unsigned char *tabs_base, int first_light_table,
                                           const bool (&hit)[2], const int (&idx)[2], const d3 (&o64)[2], const d3 (&d64v)[2],
                                           const double (&t64)[2], const unsigned (&pix)[2], float (&wt)[2], float (&cr)[2],
                                           float (&cg)[2], float (&cb)[2], bool (&final_)[2], bool (&cont)[2], RayRec (&rec)[2],
                                           int level, Counters &cnt, int &n_fp64) {
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
    cont[r] = false;
    if (hit[r]) {
      const HitGeom hg = hit_geometry(a.r.sph64, idx[r], o64[r], d64v[r], t64[r]);   // src/main.cpp:32,35
      p[r] = hg.p; n[r] = hg.n;
      m[r] = __ldg(&a.r.mat[idx[r]]); mx[r] = __ldg(&a.r.matx[idx[r]]);
      nx[r] = (float)n[r].x; ny[r] = (float)n[r].y; nz[r] = (float)n[r].z;
      // view_dir = normalized(origin - hit) = -d up to rounding (src/main.cpp:38); colour only
      vx[r] = -(float)d64v[r].x; vy[r] = -(float)d64v[r].y; vz[r] = -(float)d64v[r].z;
      sr[r] = g_frame.ambient[0] * m[r].x; sg[r] = g_frame.ambient[1] * m[r].y; sb[r] = g_frame.ambient[2] * m[r].z;
    }
  }
  const int L = a.L;
  if (__any_sync(kFull, hit[0] || hit[1])) {
    for (int l = 0; l < L; l++) {
      float dx[2], dy[2], dz[2], so[2], cosl[2];
      bool occ[2];
      const d3 lp = ldc3(g_frame.light_pos[l]);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        dx[r] = dy[r] = dz[r] = 0.f; so[r] = 0.f; cosl[r] = 0.f;
        if (hit[r]) {
          // direction light -> point: FP64 difference, FP32 normalisation (error <= 12u, see filter_math.cuh)
          const d3 w = rtx::sub(p[r], lp);
          const float wx = (float)w.x, wy = (float)w.y, wz = (float)w.z;
          const float l2 = fmaf(wz, wz, fmaf(wy, wy, wx * wx));
          const float inv = rsqrtf(l2);
          dx[r] = wx * inv; dy[r] = wy * inv; dz[r] = wz * inv;
          so[r] = l2 * inv - kEps;
          cosl[r] = -(nx[r] * dx[r] + ny[r] * dy[r] + nz[r] * dz[r]);         // n . light_dir
        }
      }
      const double *const pp[2] = {&p[0].x, &p[1].x};
      shadow_light(tab_at(tabs_base, a, first_light_table + l), a.npairs, l, dx, dy, dz, so, hit, idx, cosl, pp, a.d64,
                   a.r.sph64, occ, n_fp64);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (!hit[r]) continue;
        cnt.shadow++;
        if (occ[r]) { cnt.occluded++; if (l < 32) smask[r] |= 1u << l; continue; }
        // include/scene.h:104-117 in FP32; light_dir = -(dx,dy,dz)
        const float ndl = fmaxf(0.0f, cosl[r]);
        const float kd = (1.0f - m[r].w) * ndl;
        const float dn = -cosl[r];                                                // dot(-light_dir, n)
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
    if (!hit[r]) continue;
    if (a.r.shadow_mask) a.r.shadow_mask[(size_t)pix[r] * a.r.max_depth + level] = smask[r];
    if (mx[r].y > 0.5f) {                         // reflectivity > 0, decided in double on the host
      const float refl = m[r].w, k = wt[r] * (1.0f - refl);
      cr[r] += k * sr[r]; cg[r] += k * sg[r]; cb[r] += k * sb[r];
      wt[r] *= refl;
      if (level + 1 < a.r.max_depth) {
        reflected_ray(d64v[r], p[r], n[r], &rec[r]);
        rec[r].pix = pix[r]; rec[r].wt = wt[r]; rec[r].ar = cr[r]; rec[r].ag = cg[r]; rec[r].ab = cb[r]; rec[r].pad = 0;
        cont[r] = true;
      } else {
        final_[r] = true;                         // depth exhausted: the child contributes black
      }
    } else {
      cr[r] += wt[r] * sr[r]; cg[r] += wt[r] * sg[r]; cb[r] += wt[r] * sb[r];
      final_[r] = true;
    }
  }
}

__device__ __forceinline__ void flush_counters(const FastArgs &a, Counters &c, int &n_fp64, int level) {
  if (!a.r.counters) return;
  c.fp64 += (unsigned long long)n_fp64;
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
  c.closest = c.hits = c.shadow = c.occluded = c.fp64 = c.violations = 0;
  n_fp64 = 0;
}

// Exact FP64 t of the winner (or the brute-force safety net if the filter contradicted itself).
__device__ __forceinline__ void finish_closest(const FastArgs &a, const Best &b, const ExactRay &e, bool &hit, int &idx, double &t64,
                                               Counters &cnt, int &n_fp64) {
  double t = b.t;
  bool ok = b.exact;
  if (!ok) { n_fp64++; ok = exact_sphere(a.r.sph64, b.idx, e.o, e.d, e.a, t) && t < 1e20; }
  int bi = b.idx;
  if (!ok) { cnt.violations++; bi = exact_bruteforce(a.r.sph64, a.N, e.o, e.d, e.a, t); }
  if (bi >= 0) { hit = true; idx = bi; t64 = t; cnt.hits++; }
}

// ---------------------------------------------------------------------------------------------
// LEVEL 0: camera rays.  Each warp pulls 16x4-pixel tiles; lane = (lx, ly) = (lane & 15, lane >> 4)
// owns the two vertically adjacent pixels (x, 2*ly) and (x, 2*ly + 1) of the tile.
template <bool kSmem>
__global__ void __launch_bounds__(kThreads, 2) k_primary(const FastArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned char *tabs = a.tabs;
  if (kSmem) { stage_tables(smem, a.tabs, a.stage_bytes); tabs = smem + kSmemHeader; }
  const Tab cam = tab_at(tabs, a, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *s_rgb = smem + 64 + warp * (kWTileH * kWTileW * 3);
  const int W = a.r.W, rows = a.r.bands.local_rows, depth = a.r.max_depth;
  Counters cnt = {0, 0, 0, 0, 0, 0};
  int n_fp64 = 0;
  for (;;) {
    const int tile = warp_fetch(a.tile_counter);
    if (tile >= a.nwtiles) break;
    const int tx0 = (tile % a.wtiles_x) * kWTileW, ty0 = (tile / a.wtiles_x) * kWTileH;
    const int x = tx0 + (lane & 15);
    int lr[2], j[2];
    unsigned pix[2];
    bool live[2];
    float dx[2], dy[2], dz[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      lr[r] = ty0 + (lane >> 4) * 2 + r;
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
    const RaySrc src[2] = {{a.r.su, a.r.sv, x, j[0], nullptr}, {a.r.su, a.r.sv, x, j[1], nullptr}};
    closest_shared(cam, a.npairs, dx, dy, dz, live, a.d64, a.r.sph64, src, best);

    bool hit[2], final_[2], cont[2];
    int idx[2];
    d3 o64[2], d64v[2];
    double t64[2];
    RayRec rec[2];
    float wt[2] = {1.f, 1.f}, cr[2] = {0.f, 0.f}, cg[2] = {0.f, 0.f}, cb[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; r++) {
      hit[r] = false; final_[r] = false; idx[r] = -1; t64[r] = 0;
      o64[r] = rtx::mk(0, 0, 0); d64v[r] = o64[r];
      if (!live[r]) continue;
      cnt.closest++;
      n_fp64 += best[r].nfp64;
      if (best[r].idx >= 0) {
        const ExactRay e = exact_ray(src[r]);
        finish_closest(a, best[r], e, hit[r], idx[r], t64[r], cnt, n_fp64);
        o64[r] = e.o; d64v[r] = e.d;
      }
      if (a.r.hit_idx) a.r.hit_idx[(size_t)pix[r] * depth] = idx[r];
      if (!hit[r]) {                               // sky, src/main.cpp:26-30
        const float ts = 0.5f * (dy[r] + 1.0f);
        cr[r] = (1.0f - ts) + 0.5f * ts; cg[r] = (1.0f - ts) + 0.7f * ts; cb[r] = (1.0f - ts) + ts;
        final_[r] = true;
      }
    }
    shade_hits(a, tabs, 1, hit, idx, o64, d64v, t64, pix, wt, cr, cg, cb, final_, cont, rec, 0, cnt, n_fp64);
#pragma unroll
    for (int r = 0; r < 2; r++) queue_push(cont[r], rec[r], a.q_out, a.q_out_count);

    // ---- 8-bit quantise (src/main.cpp:84-86) into the warp's staging buffer, then 128-bit row stores
#pragma unroll
    for (int r = 0; r < 2; r++) {
      unsigned char *q = s_rgb + (((lane >> 4) * 2 + r) * kWTileW + (lane & 15)) * 3;
      const bool fin = final_[r] || (x < W && lr[r] < rows && depth <= 0);
      q[0] = (unsigned char)quant8(fin ? cr[r] : 0.f); q[1] = (unsigned char)quant8(fin ? cg[r] : 0.f);
      q[2] = (unsigned char)quant8(fin ? cb[r] : 0.f);
    }
    __syncwarp();
    if (tx0 + kWTileW <= W && (W & 15) == 0) {
      // 3 x 16-byte stores per 48-byte row segment; pixels still in flight are overwritten by k_bounce
      if (lane < kWTileH * 3) {
        const int ty = lane / 3, seg = lane % 3;
        if (ty0 + ty < rows)
          *reinterpret_cast<uint4 *>(a.r.rgb + ((size_t)(ty0 + ty) * W + tx0) * 3 + seg * 16) =
              *reinterpret_cast<const uint4 *>(s_rgb + ty * kWTileW * 3 + seg * 16);
      }
    } else {
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (x < W && lr[r] < rows) {
          const unsigned char *q = s_rgb + (((lane >> 4) * 2 + r) * kWTileW + (lane & 15)) * 3;
          unsigned char *o = a.r.rgb + (size_t)pix[r] * 3;
          o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
        }
      }
    }
    __syncwarp();
  }
  flush_counters(a, cnt, n_fp64, 0);
}

// ---------------------------------------------------------------------------------------------
// LEVEL >= 1: reflected rays from the queue, two per lane, 64 per warp fetch.
// kTail = false: one level, survivors are compacted into q_out.
// kTail = true : every lane follows its ray to termination; the record is updated in place.
template <bool kSmem, bool kTail>
__global__ void __launch_bounds__(kThreads, 2) k_bounce(const FastArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned nq = *a.q_in_count;
  if (nq == 0u) return;                             // nothing survived to this level
  // staged: [L light tables][general table], contiguous in global memory in that order
  const unsigned char *tabs = a.tabs + a.tstride;
  if (kSmem) { stage_tables(smem, tabs, a.stage_bytes); tabs = smem + kSmemHeader; }
  const float4 *gen = reinterpret_cast<const float4 *>(tabs + (size_t)a.L * a.tstride);
  const int lane = threadIdx.x & 31;
  const int depth = a.r.max_depth;
  Counters cnt = {0, 0, 0, 0, 0, 0};
  int n_fp64 = 0;
  RayRec *qin = a.q_in;
  for (;;) {
    // level 1: 64 rays per fetch (two per lane); tail: 32 (one per lane -- it is latency bound, the
    // SMs are mostly empty there, so shorter per-warp chains beat packed arithmetic)
    constexpr unsigned kRaysPerFetch = kTail ? 32u : 64u;
    const int chunk = warp_fetch(a.chunk_counter);
    if ((unsigned)chunk * kRaysPerFetch >= nq) break;
    bool live[2];
    unsigned qi[2], pix[2];
    float ox[2], oy[2], oz[2], dx[2], dy[2], dz[2];
    float wt[2], cr[2], cg[2], cb[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      // the two rays of a lane are neighbours in the queue (coherent)
      qi[r] = kTail ? (unsigned)chunk * 32u + lane : (unsigned)chunk * 64u + 2u * lane + r;
      live[r] = qi[r] < nq && !(kTail && r == 1);
      ox[r] = oy[r] = oz[r] = dx[r] = dy[r] = dz[r] = 0.f; wt[r] = cr[r] = cg[r] = cb[r] = 0.f; pix[r] = 0;
      if (live[r]) {
        const RayRec &q = qin[qi[r]];
        // recentred FP32 origin and FP32 direction for the filter (exact values stay in the record)
        ox[r] = (float)(q.ox - a.c0[0]); oy[r] = (float)(q.oy - a.c0[1]); oz[r] = (float)(q.oz - a.c0[2]);
        dx[r] = (float)q.dx; dy[r] = (float)q.dy; dz[r] = (float)q.dz;
        pix[r] = q.pix; wt[r] = q.wt; cr[r] = q.ar; cg[r] = q.ag; cb[r] = q.ab;
      }
    }
    const RaySrc src[2] = {{nullptr, nullptr, 0, 0, qin + qi[0]}, {nullptr, nullptr, 0, 0, qin + qi[1]}};
    for (int level = a.level;; level++) {
      Best best[2];
      best_init(best[0]); best_init(best[1]);
      closest_general(gen, a.npairs, a.N, ox, oy, oz, dx, dy, dz, live, a.d64, a.gS2, a.g_dtmax, a.r.sph64, src, best);
      bool hit[2], final_[2], cont[2];
      int idx[2];
      d3 o64[2], d64v[2];
      double t64[2];
      RayRec rec[2];
#pragma unroll
      for (int r = 0; r < 2; r++) {
        hit[r] = false; final_[r] = false; idx[r] = -1; t64[r] = 0;
        o64[r] = rtx::mk(0, 0, 0); d64v[r] = o64[r];
        if (!live[r]) continue;
        cnt.closest++;
        n_fp64 += best[r].nfp64;
        if (best[r].idx >= 0) {
          const ExactRay e = exact_ray(src[r]);
          finish_closest(a, best[r], e, hit[r], idx[r], t64[r], cnt, n_fp64);
          o64[r] = e.o; d64v[r] = e.d;
        }
        if (a.r.hit_idx) a.r.hit_idx[(size_t)pix[r] * depth + level] = idx[r];
        if (!hit[r]) {
          const float ts = 0.5f * (dy[r] + 1.0f);
          cr[r] += wt[r] * ((1.0f - ts) + 0.5f * ts); cg[r] += wt[r] * ((1.0f - ts) + 0.7f * ts); cb[r] += wt[r] * ((1.0f - ts) + ts);
          final_[r] = true;
        }
      }
      shade_hits(a, tabs, 0, hit, idx, o64, d64v, t64, pix, wt, cr, cg, cb, final_, cont, rec, level, cnt, n_fp64);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (live[r] && final_[r]) {
          unsigned char *o = a.r.rgb + (size_t)pix[r] * 3;
          o[0] = (unsigned char)quant8(cr[r]); o[1] = (unsigned char)quant8(cg[r]); o[2] = (unsigned char)quant8(cb[r]);
        }
      }
      if (!kTail) {
#pragma unroll
        for (int r = 0; r < 2; r++) queue_push(cont[r], rec[r], a.q_out, a.q_out_count);
        break;
      }
      if (a.r.counters) flush_counters(a, cnt, n_fp64, level);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        live[r] = cont[r];
        if (cont[r]) {
          qin[qi[r]] = rec[r];                    // in place: the exact ray is re-read from here
          ox[r] = (float)(rec[r].ox - a.c0[0]); oy[r] = (float)(rec[r].oy - a.c0[1]); oz[r] = (float)(rec[r].oz - a.c0[2]);
          dx[r] = (float)rec[r].dx; dy[r] = (float)rec[r].dy; dz[r] = (float)rec[r].dz;
        }
      }
      if (!__any_sync(kFull, live[0] || live[1])) break;
    }
  }
  if (!kTail) flush_counters(a, cnt, n_fp64, a.level);
}

}  // namespace rtf
#endif
