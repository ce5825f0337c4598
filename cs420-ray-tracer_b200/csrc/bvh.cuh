// bvh.cuh -- device-built LBVH over the spheres (large scenes: BASELINE config 5) and its traversal.
//
// BUILD (all on the device, once per rt_upload_scene; Karras 2012 "Maximizing parallelism in the
// construction of BVHs"):
//   k_bvh_bounds   AABB of the recentred FP32 sphere centres            (atomics on order-preserving ints)
//   k_bvh_morton   30-bit Morton code of every centre
//   k_bvh_sort     stable LSD radix sort of (code, sphere), 4-bit digits, one CTA: per-thread contiguous
//                  chunks, a [digit][thread] count table in shared memory, one block scan per pass
//   k_bvh_leaves   every sorted sphere becomes a leaf (box = the sphere's box, inflated by `eps`); the
//                  benchmark scenes are sparse, so wider leaves would mostly enclose empty space
//   k_bvh_hier     one thread per internal node: range + split from common-prefix lengths (ties on the
//                  code are broken by the leaf index, so duplicate codes are fine)
//   k_bvh_refit    bottom-up box union; the second thread to arrive at a node does the work
//   k_bvh_pack     final 64-byte nodes that carry BOTH children's boxes (one fetch decides both)
//
// TRAVERSAL is per lane (its own ray, its own stack in local memory), near child first.  It only
// selects CANDIDATE spheres: every candidate goes through the same conservative FP32 bracketing and
// exact FP64 deciders as the brute-force paths, so hit indices and shadow booleans stay bit-exact.
// The slab test is conservative: boxes are inflated by `eps` (>= 64u x 3S: FP32 rounding of recentred
// centres, ray origins and the 12u direction error over any distance <= 2S inside the scene ball), the
// computed entry/exit distances are widened by a relative 1e-6 (their FP32 error is <= 3u).
#ifndef RT_BVH_CUH
#define RT_BVH_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace rtb {

constexpr int kLeafSize = 1;
constexpr int kStack = 64;        // >= 30 code bits + 32 index bits of the deepest possible LBVH

struct __align__(16) BvhNode {    // 64 bytes
  float4 a;                       // lo0.x lo0.y lo0.z hi0.x
  float4 b;                       // hi0.y hi0.z lo1.x lo1.y
  float4 c;                       // lo1.z hi1.x hi1.y hi1.z
  int4 d;                         // child0, child1 (>= 0 internal node, < 0: leaf ~index), -, -
};

struct BvhView {
  const BvhNode *nodes;           // nleaf - 1 internal nodes (at least one: see k_bvh_pack), root = 0; a child < 0 is
                                  // the leaf ~child, which after k_bvh_pack IS the sphere index (kLeafSize == 1)
};

// ---------------------------------------------------------------------------------------------
// build
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// cen[i] = (cx', cy', cz', r) recentred FP32.  bounds[0..2] = min, [3..5] = max (order-preserving ints)
__global__ void k_bvh_bounds(const float4 *cen, int n, int *bounds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
  for (; i < n; i += gridDim.x * blockDim.x) {
    const float4 c = cen[i];
    lo[0] = fminf(lo[0], c.x); lo[1] = fminf(lo[1], c.y); lo[2] = fminf(lo[2], c.z);
    hi[0] = fmaxf(hi[0], c.x); hi[1] = fmaxf(hi[1], c.y); hi[2] = fmaxf(hi[2], c.z);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&bounds[k], f2ord(lo[k])); atomicMax(&bounds[3 + k], f2ord(hi[k])); }
  }
}

__device__ __forceinline__ unsigned expand10(unsigned v) {   // 10 bits -> every third bit
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

__global__ void k_bvh_morton(const float4 *cen, int n, const int *bounds, unsigned *keys, int *vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 c = cen[i];
  const float p[3] = {c.x, c.y, c.z};
  unsigned q[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float lo = ord2f(bounds[k]), hi = ord2f(bounds[3 + k]);
    const float ext = fmaxf(hi - lo, 1e-30f);
    q[k] = (unsigned)fminf(fmaxf((p[k] - lo) / ext * 1024.0f, 0.0f), 1023.0f);
  }
  keys[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
  vals[i] = i;
}

// Stable LSD radix sort, one CTA of 1024 threads, 4-bit digits; `passes` x 4 bits are sorted.
// Result ends in (keys_a, vals_a) when `passes` is even.
constexpr int kSortThreads = 1024;
__global__ void __launch_bounds__(kSortThreads, 1) k_bvh_sort(unsigned *keys_a, int *vals_a, unsigned *keys_b, int *vals_b, int n, int passes) {
  extern __shared__ unsigned s_cnt[];                 // [16][1024] counts, then offsets; + 32 warp sums
  unsigned *s_warp = s_cnt + 16 * kSortThreads;
  const int t = threadIdx.x;
  const int chunk = (n + kSortThreads - 1) / kSortThreads;
  const int beg = min(t * chunk, n), end = min(beg + chunk, n);
  unsigned *ks = keys_a, *kd = keys_b;
  int *vs = vals_a, *vd = vals_b;
  for (int pass = 0; pass < passes; pass++) {
    const int shift = pass * 4;
    for (int d = 0; d < 16; d++) s_cnt[d * kSortThreads + t] = 0u;
    for (int i = beg; i < end; i++) s_cnt[((ks[i] >> shift) & 15u) * kSortThreads + t]++;
    __syncthreads();
    // exclusive scan of the flattened [digit][thread] table: thread t owns entries 16t .. 16t+15
    unsigned loc[16], sum = 0u;
#pragma unroll
    for (int k = 0; k < 16; k++) { loc[k] = sum; sum += s_cnt[t * 16 + k]; }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= o) incl += v; }
    if ((t & 31) == 31) s_warp[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
      unsigned w = s_warp[t], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { unsigned v = __shfl_up_sync(0xffffffffu, wi, o); if (t >= o) wi += v; }
      s_warp[t] = wi - w;
    }
    __syncthreads();
    const unsigned base = s_warp[t >> 5] + incl - sum;
#pragma unroll
    for (int k = 0; k < 16; k++) s_cnt[t * 16 + k] = base + loc[k];
    __syncthreads();
    for (int i = beg; i < end; i++) {
      const unsigned key = ks[i];
      const unsigned pos = s_cnt[((key >> shift) & 15u) * kSortThreads + t]++;
      kd[pos] = key; vd[pos] = vs[i];
    }
    __syncthreads();
    unsigned *tk = ks; ks = kd; kd = tk;
    int *tv = vs; vs = vd; vd = tv;
  }
}

// leaf l = sorted positions [l*kLeafSize, (l+1)*kLeafSize): box (inflated by eps), key of its first sphere
__global__ void k_bvh_leaves(const float4 *cen, const unsigned *keys, const int *vals, int n, int nleaf, float eps,
                             float4 *leaf_lo, float4 *leaf_hi, unsigned *leaf_key) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nleaf) return;
  int id[kLeafSize];
  float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
#pragma unroll
  for (int k = 0; k < kLeafSize; k++) {
    const int p = l * kLeafSize + k;
    id[k] = p < n ? vals[p] : -1;
    if (id[k] >= 0) {
      const float4 c = cen[id[k]];
      const float r = c.w + eps;
      lo[0] = fminf(lo[0], __fsub_rd(c.x, r)); lo[1] = fminf(lo[1], __fsub_rd(c.y, r)); lo[2] = fminf(lo[2], __fsub_rd(c.z, r));
      hi[0] = fmaxf(hi[0], __fadd_ru(c.x, r)); hi[1] = fmaxf(hi[1], __fadd_ru(c.y, r)); hi[2] = fmaxf(hi[2], __fadd_ru(c.z, r));
    }
  }
  leaf_lo[l] = make_float4(lo[0], lo[1], lo[2], 0.f);
  leaf_hi[l] = make_float4(hi[0], hi[1], hi[2], 0.f);
  leaf_key[l] = keys[l * kLeafSize];
}

__device__ __forceinline__ int bvh_delta(const unsigned *key, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned a = key[i], b = key[j];
  return a == b ? 32 + __clz((unsigned)i ^ (unsigned)j) : __clz(a ^ b);
}

// internal node i of n-1; children: >= 0 internal, < 0 leaf (~leaf)
__global__ void k_bvh_hier(const unsigned *key, int n, int *left, int *right, int *parent_int, int *parent_leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = bvh_delta(key, n, i, i + 1) - bvh_delta(key, n, i, i - 1) >= 0 ? 1 : -1;
  const int dmin = bvh_delta(key, n, i, i - d);
  int lmax = 2;
  while (bvh_delta(key, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (bvh_delta(key, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = bvh_delta(key, n, i, j);
  int s = 0;
  for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
    if (bvh_delta(key, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int lc = (lo == gamma) ? ~gamma : gamma;
  const int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  left[i] = lc; right[i] = rc;
  if (lc >= 0) parent_int[lc] = i; else parent_leaf[~lc] = i;
  if (rc >= 0) parent_int[rc] = i; else parent_leaf[~rc] = i;
  if (i == 0) parent_int[0] = -1;
}

__global__ void k_bvh_refit(int nleaf, const int *left, const int *right, const int *parent_int, const int *parent_leaf,
                            const float4 *leaf_lo, const float4 *leaf_hi, float4 *node_lo, float4 *node_hi, int *flags) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nleaf) return;
  int node = parent_leaf[l];
  while (node >= 0) {
    __threadfence();
    if (atomicAdd(&flags[node], 1) == 0) return;     // the sibling subtree is not finished: its thread continues
    __threadfence();
    const int lc = left[node], rc = right[node];
    const volatile float4 *llo = lc < 0 ? &leaf_lo[~lc] : &node_lo[lc], *lhi = lc < 0 ? &leaf_hi[~lc] : &node_hi[lc];
    const volatile float4 *rlo = rc < 0 ? &leaf_lo[~rc] : &node_lo[rc], *rhi = rc < 0 ? &leaf_hi[~rc] : &node_hi[rc];
    node_lo[node] = make_float4(fminf(llo->x, rlo->x), fminf(llo->y, rlo->y), fminf(llo->z, rlo->z), 0.f);
    node_hi[node] = make_float4(fmaxf(lhi->x, rhi->x), fmaxf(lhi->y, rhi->y), fmaxf(lhi->z, rhi->z), 0.f);
    node = parent_int[node];
  }
}

// nint = max(nleaf - 1, 1) packed nodes.  A single-leaf tree gets one node whose second child is empty.
__global__ void k_bvh_pack(int nleaf, const int *left, const int *right, const int *vals, const int *orig, const float4 *leaf_lo, const float4 *leaf_hi,
                           const float4 *node_lo, const float4 *node_hi, BvhNode *out) {
  static_assert(kLeafSize == 1, "leaf children are rewritten to sphere indices");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max(nleaf - 1, 1)) return;
  float4 l0, h0, l1, h1;
  int c0, c1;
  if (nleaf == 1) {
    c0 = ~0; c1 = ~0;
    l0 = leaf_lo[0]; h0 = leaf_hi[0];
    l1 = make_float4(3e38f, 3e38f, 3e38f, 0.f); h1 = make_float4(-3e38f, -3e38f, -3e38f, 0.f);
  } else {
    c0 = left[i]; c1 = right[i];
    l0 = c0 < 0 ? leaf_lo[~c0] : node_lo[c0]; h0 = c0 < 0 ? leaf_hi[~c0] : node_hi[c0];
    l1 = c1 < 0 ? leaf_lo[~c1] : node_lo[c1]; h1 = c1 < 0 ? leaf_hi[~c1] : node_hi[c1];
  }
  BvhNode nd;
  nd.a = make_float4(l0.x, l0.y, l0.z, h0.x);
  nd.b = make_float4(h0.y, h0.z, l1.x, l1.y);
  nd.c = make_float4(l1.z, h1.x, h1.y, h1.z);
  nd.d = make_int4(c0 < 0 ? ~orig[vals[~c0]] : c0, c1 < 0 ? ~orig[vals[~c1]] : c1, 0, 0);   // orig: build order -> sphere index
  out[i] = nd;
}

// ---------------------------------------------------------------------------------------------
// traversal
struct BvhRay {
  float ox, oy, oz, ix, iy, iz;   // origin, reciprocal direction (components clamped away from 0)
};
__device__ __forceinline__ float clamp_dir(float d) { return fabsf(d) < 1e-30f ? copysignf(1e-30f, d) : d; }
__device__ __forceinline__ BvhRay bvh_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
  BvhRay r;
  r.ox = ox; r.oy = oy; r.oz = oz;
  r.ix = __frcp_rn(clamp_dir(dx)); r.iy = __frcp_rn(clamp_dir(dy)); r.iz = __frcp_rn(clamp_dir(dz));
  return r;
}
// widened entry distance of the box, or 3e38 when the ray segment [t0, t1] misses it
__device__ __forceinline__ float bvh_slab(const BvhRay &r, float lx, float ly, float lz, float hx, float hy, float hz, float t0, float t1) {
  const float ax = (lx - r.ox) * r.ix, bx = (hx - r.ox) * r.ix;
  const float ay = (ly - r.oy) * r.iy, by = (hy - r.oy) * r.iy;
  const float az = (lz - r.oz) * r.iz, bz = (hz - r.oz) * r.iz;
  float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
  float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
  tn = tn - fabsf(tn) * 1e-6f;
  tf = tf + fabsf(tf) * 1e-6f;
  return (tn <= tf && tn <= t1 && tf >= t0) ? tn : 3.0e38f;
}

// Calls leaf(sphere_index) for every sphere whose (inflated) box the segment [t0, t1] of the ray touches; leaf()
// returns the (possibly smaller) t1 to continue with, or a value < t0 to stop the traversal.  "while-while":
// the lanes of a warp first all descend to their next leaf, then all run the (long, branchy) leaf code together.
template <typename F>
__device__ __forceinline__ void bvh_traverse(const BvhView v, const BvhRay &r, float t0, float t1, F &&leaf) {
  int stack[kStack];
  int sp = 0, node = 0;
  constexpr int kDone = 0x7fffffff;
  for (;;) {
    while (node >= 0 && node != kDone) {
      const BvhNode *nd = v.nodes + node;
      const float4 a = __ldg(&nd->a), b = __ldg(&nd->b), c = __ldg(&nd->c);
      const int4 d = __ldg(&nd->d);
      const float e0 = bvh_slab(r, a.x, a.y, a.z, a.w, b.x, b.y, t0, t1);
      const float e1 = bvh_slab(r, b.z, b.w, c.x, c.y, c.z, c.w, t0, t1);
      const bool h0 = e0 < 3.0e38f, h1 = e1 < 3.0e38f;
      if (h0 && h1) {
        const bool swap = e1 < e0;
        stack[sp++] = swap ? d.x : d.y;             // far child later
        node = swap ? d.y : d.x;
      } else if (h0 || h1) {
        node = h0 ? d.x : d.y;
      } else {
        node = sp > 0 ? stack[--sp] : kDone;
      }
    }
    if (node == kDone) return;
    t1 = leaf(~node);
    if (t1 < t0) return;
    node = sp > 0 ? stack[--sp] : kDone;
  }
}

// Resumable form of the same traversal, for kernels that keep every lane busy by handing it a new ray as soon as its
// old one is finished (persistent threads with dynamic ray fetch): bvh_next() advances to the next candidate sphere.
struct BvhIter {
  int stack[kStack];
  int sp, node;
  float t0, t1;
  BvhRay r;
};
constexpr int kBvhDone = 0x7fffffff;
__device__ __forceinline__ void bvh_begin(BvhIter &it, const BvhRay &r, float t0, float t1) {
  it.sp = 0; it.node = 0; it.t0 = t0; it.t1 = t1; it.r = r;
}
// Visits at most `budget` internal nodes.  Returns the next candidate sphere (>= 0), -1 when the traversal is complete,
// -2 when the budget ran out first (call again).
__device__ __forceinline__ int bvh_next(const BvhView v, BvhIter &it, int budget) {
  int node = it.node, sp = it.sp;
  while (node >= 0 && node != kBvhDone && budget-- > 0) {
    const BvhNode *nd = v.nodes + node;
    const float4 a = __ldg(&nd->a), b = __ldg(&nd->b), c = __ldg(&nd->c);
    const int4 d = __ldg(&nd->d);
    const float e0 = bvh_slab(it.r, a.x, a.y, a.z, a.w, b.x, b.y, it.t0, it.t1);
    const float e1 = bvh_slab(it.r, b.z, b.w, c.x, c.y, c.z, c.w, it.t0, it.t1);
    const bool h0 = e0 < 3.0e38f, h1 = e1 < 3.0e38f;
    if (h0 && h1) {
      const bool swap = e1 < e0;
      it.stack[sp++] = swap ? d.x : d.y;
      node = swap ? d.y : d.x;
    } else if (h0 || h1) {
      node = h0 ? d.x : d.y;
    } else {
      node = sp > 0 ? it.stack[--sp] : kBvhDone;
    }
  }
  int ret;
  if (node == kBvhDone) ret = -1;
  else if (node >= 0) ret = -2;
  else { ret = ~node; node = sp > 0 ? it.stack[--sp] : kBvhDone; }
  it.node = node; it.sp = sp;
  return ret;
}

}  // namespace rtb
#endif
