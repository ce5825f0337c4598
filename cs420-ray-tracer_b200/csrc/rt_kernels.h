// rt_kernels.h -- host-callable launchers implemented in rt_kernels.cu.
#ifndef RT_KERNELS_H
#define RT_KERNELS_H

#include <cuda_runtime.h>

#include "rt_device.h"

// Copies the frame constants into this device's __constant__ bank (async on `stream`).
cudaError_t rtk_set_frame_const(const RtFrameConst *host_const, cudaStream_t stream);

// mode 1: exact FP64 brute force.  Returns the number of kernels launched (or <0: cudaError).
int rtk_launch_exact(const RtRenderArgs &args, cudaStream_t stream);


// mode 0: FP32-filter fast path (kernels_fast.cuh).
struct RtFastScene {
  int N, L, npairs;
  void *cam32;      // npairs x 2 float4 : camera-origin table
  void *light32;    // L x npairs x 2 float4 : per-light tables
  void *sph32;      // npairs x 3 float4 : general-origin table
};
struct RtFastWork {
  int num_sms;
  void *queue[2];   // secondary-ray records, ping-pong
  size_t queue_cap;
  unsigned int *qcount;  // device: [2] record counts
  void *accum;      // reserved
  size_t accum_cap;
};
int rtk_fast_init(int device);
int rtk_fast_build_scene(RtFastScene *fs, const double *spheres, int N, const RtFrameConst *frame, cudaStream_t stream);
void rtk_fast_free_scene(RtFastScene *fs);
void rtk_fast_free_work(RtFastWork *w);
int rtk_launch_fast(const RtRenderArgs &args, const RtFastScene *fs, RtFastWork *w, cudaStream_t stream);

#endif
