// rt_kernels.cu -- the single device translation unit of librt_b200.so (sm_100a only).
#include "rt_kernels.h"

// Camera basis, lights and ambient: one copy per device, refreshed by rt_upload_scene.
__constant__ RtFrameConst g_frame;

#include "kernels_exact.cuh"

cudaError_t rtk_set_frame_const(const RtFrameConst *host_const, cudaStream_t stream) {
  return cudaMemcpyToSymbolAsync(g_frame, host_const, sizeof(RtFrameConst), 0, cudaMemcpyHostToDevice, stream);
}

int rtk_launch_exact(const RtRenderArgs &args, cudaStream_t stream) {
  dim3 block(128);
  dim3 grid((args.W + 31) / 32, (args.bands.local_rows + 3) / 4);
  rtk::k_exact<<<grid, block, 0, stream>>>(args);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}

// ---- fast path: placeholder until kernels_fast.cuh lands (fails loudly, never falls back) ----
int rtk_fast_init(int) { return 0; }
int rtk_fast_build_scene(RtFastScene *fs, const double *, int N, const RtFrameConst *f, cudaStream_t) { fs->N = N; fs->L = f->nlights; return 0; }
void rtk_fast_free_scene(RtFastScene *) {}
void rtk_fast_free_work(RtFastWork *) {}
int rtk_launch_fast(const RtRenderArgs &, const RtFastScene *, RtFastWork *, cudaStream_t) { return -(int)cudaErrorNotSupported; }
