// kernels_frame.cuh -- the WHOLE FRAME in one persistent kernel (scenes whose tables fit in shared memory).
//
// The wavefront of kernels_wave.cuh is three launches per reflection level plus a tail; on the headline frame
// (complex.txt, 1920x1080, depth 5: 0.6 ms) the launches, their ramp-up, the table staging of every launch and -- above
// all -- the level-serial tail (one warp's dependency chain per level, 9.7 % occupancy) were 45 % of the frame time
// for 13 % of its rays.  Here ONE cooperative grid (2 CTAs per SM, all co-resident) stages every table of the scene
// once -- camera, L lights, general: ONE TMA bulk copy per CTA -- and then walks the SAME phases
//
//     level 0: closest0 | barrier | shadow + shade | barrier | level 1: closest1 | barrier | shadow + shade | ...
//
// with a grid barrier (one atomic arrive + spin per CTA, ~2 us) where the wavefront had a kernel boundary.  Every
// phase is the unchanged body of the corresponding wavefront kernel (closest0_body / closest1_body / shadow_body),
// fed from the same atomic work counters, so all warps of the machine share the rays of EVERY level -- also of the
// deep ones that the tail kernel used to walk one warp per 32 rays.  Shading is fused into the shadow phase (the warp
// that learns the last occlusion bit of a chunk of hits shades it), so the hit records are read once.
#ifndef RT_KERNELS_FRAME_CUH
#define RT_KERNELS_FRAME_CUH

#include "kernels_wave.cuh"

#ifndef RT_FRAME_CTAS
#define RT_FRAME_CTAS 2
#endif

namespace rtf {

// Grid-wide barrier of a cooperative launch: every CTA arrives once per call; `target` (per thread, only thread 0's
// counts) is the arrival total this call waits for.  Release / acquire through the fences around the arrive + spin;
// thread 0's acquire fence also drops the SM's L1 lines, so the plain loads that follow see what other SMs wrote.
// A CTA that waits longer than a second gives up, flags CTL_ERR and lets the kernel run out (wrong pixels, but no hang).
__device__ __forceinline__ void grid_barrier(unsigned int *ctl, unsigned &target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(ctl + CTL_BARRIER, 1u);
    volatile unsigned int *bar = ctl + CTL_BARRIER;
    if (*bar < target) {
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
      while (*bar < target) {
        __nanosleep(32);
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > 1000000000ULL) { atomicOr(ctl + CTL_ERR, 4u); break; }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// phase time stamps for diagnostics (RT_FRAME_TRACE=1 prints them after a stats render): ctl[CTL_TRACE + k]
constexpr int CTL_TRACE = 242;
__device__ __forceinline__ void trace_stamp(unsigned int *ctl, int &k) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && k < 14) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    ctl[CTL_TRACE + k] = (unsigned)t;
  }
  k++;
}

template <bool kFuse>
__global__ void __launch_bounds__(kThreads, RT_FRAME_CTAS) k_frame(const WaveArgs w) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FastArgs &a = w.f;
  int tk = 0;
  trace_stamp(w.ctl, tk);
  // staged once: (1 + L) shared-origin tables and the general table, contiguous in global memory in that order
  stage_tables(smem, a.tabs, a.stage_bytes);
  const unsigned char *tabs = smem + kSmemHeader;
  unsigned char *wbase = smem + kSmemHeader + ((a.stage_bytes + 127u) & ~127u);
  const float4 *gen = reinterpret_cast<const float4 *>(tabs + (size_t)(a.L + 1) * a.tstride);
  // the control words of the NEXT frame: nobody looks at them before this kernel has exited
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < CTL_WORDS; i += blockDim.x) w.ctl_next[i] = 0u;
  unsigned target = 0;
  const int depth = a.r.max_depth;
  for (int level = 0; level < depth; level++) {
    const Lvl lv = lvl_at(w, level);
    if (level == 0) closest0_body<kTabSmem>(w, lv, smem, tabs, wbase, !kFuse && tile_path(a));   // (fused shading writes pixel by pixel)
    else closest1_body<false>(w, lv, gen);
    grid_barrier(w.ctl, target);
    trace_stamp(w.ctl, tk);
    if (__ldcg(lv.hit_count) == 0u) break;             // (every CTA reads the same final count: a uniform decision)
    shadow_body<kTabSmem, kFuse>(w, lv, smem, tabs + a.tstride, wbase);
    if (!kFuse) {
      grid_barrier(w.ctl, target);
      trace_stamp(w.ctl, tk);
      shade_body(w, lv);
    }
    if (level + 1 >= depth) break;
    grid_barrier(w.ctl, target);
    trace_stamp(w.ctl, tk);
    if (__ldcg(lv.q_out_count) == 0u) break;
  }
  trace_stamp(w.ctl, tk);
}

}  // namespace rtf
#endif
