// microbench.cu -- measures the denominators this project's roofline needs on the box it runs on:
// SM count, SM clock under load, FP32 FFMA / FFMA2 and FP64 DFMA issue peaks, and the raw rate of
// the filter inner loop variants (scalar FFMA vs packed FFMA2 over sphere pairs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench scripts/microbench.cu
// Prints one JSON object.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ---- pure issue peaks -------------------------------------------------------------------------
template <int ILP>
__global__ void k_ffma(float *out, int iters, float a, float b, unsigned long long *clk) {
  float x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 0.001f + i;
  unsigned long long c0 = clock64(), g0 = gtime();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 16; k++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
  }
  unsigned long long c1 = clock64(), g1 = gtime();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = c1 - c0; clk[1] = g1 - g0; }
}

template <int ILP>
__global__ void k_ffma2(float2 *out, int iters, float2 a, float2 b) {
  float2 x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 16; k++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = __ffma2_rn(x[i], a, b);
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < ILP; i++) { s.x += x[i].x; s.y += x[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 0.001 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 16; k++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- filter inner loop prototypes -------------------------------------------------------------
// scalar: table entry per sphere = (ocx, ocy, ocz, -cc'); D = (oc.d)^2 - cc'
template <int R>
__global__ void __launch_bounds__(256) k_filter_scalar(const float4 *tab, int n, int iters, unsigned *out) {
  extern __shared__ float4 s[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = tab[i];
  __syncthreads();
  float dx[R], dy[R], dz[R];
  unsigned acc[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    float u = (threadIdx.x * R + r) * 1e-4f + blockIdx.x * 1e-3f;
    dx[r] = u; dy[r] = 0.3f - u; dz[r] = -0.9f; acc[r] = 0xffffffffu;
  }
  unsigned slow = 0;
  for (int it = 0; it < iters; it++) {
    for (int i = 0; i < n; i += 8) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        float4 a = s[i + k];
#pragma unroll
        for (int r = 0; r < R; r++) {
          float t = a.x * dx[r];
          t = fmaf(a.y, dy[r], t);
          t = fmaf(a.z, dz[r], t);
          float D = fmaf(t, t, a.w);
          acc[r] &= __float_as_uint(D);
        }
      }
      unsigned m = acc[0];
#pragma unroll
      for (int r = 1; r < R; r++) m &= acc[r];
      if ((int)m >= 0) {  // some D >= 0: slow path
        slow++;
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0xffffffffu;
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) dx[r] += 1e-7f;
  }
  unsigned o = slow;
#pragma unroll
  for (int r = 0; r < R; r++) o += acc[r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = o;
}

// packed: table entry per sphere PAIR = {(ocx0, ocx1, ocy0, ocy1), (ocz0, ocz1, -cc0, -cc1)}
template <int R>
__global__ void __launch_bounds__(256) k_filter_pairs(const float4 *tab, int npairs, int iters, unsigned *out) {
  extern __shared__ float4 s[];
  for (int i = threadIdx.x; i < 2 * npairs; i += blockDim.x) s[i] = tab[i];
  __syncthreads();
  float2 dx[R], dy[R], dz[R];
  unsigned acc[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    float u = (threadIdx.x * R + r) * 1e-4f + blockIdx.x * 1e-3f;
    dx[r] = make_float2(u, u); dy[r] = make_float2(0.3f - u, 0.3f - u); dz[r] = make_float2(-0.9f, -0.9f); acc[r] = 0xffffffffu;
  }
  unsigned slow = 0;
  for (int it = 0; it < iters; it++) {
    for (int i = 0; i < npairs; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float4 a = s[2 * (i + k)], b = s[2 * (i + k) + 1];
#pragma unroll
        for (int r = 0; r < R; r++) {
          float2 t = __fmul2_rn(make_float2(a.x, a.y), dx[r]);
          t = __ffma2_rn(make_float2(a.z, a.w), dy[r], t);
          t = __ffma2_rn(make_float2(b.x, b.y), dz[r], t);
          float2 D = __ffma2_rn(t, t, make_float2(b.z, b.w));
          acc[r] &= __float_as_uint(D.x) & __float_as_uint(D.y);
        }
      }
      unsigned m = acc[0];
#pragma unroll
      for (int r = 1; r < R; r++) m &= acc[r];
      if ((int)m >= 0) {
        slow++;
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0xffffffffu;
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) dx[r].x += 1e-7f;
  }
  unsigned o = slow;
#pragma unroll
  for (int r = 0; r < R; r++) o += acc[r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = o;
}

// general-origin packed: per pair {(cx0,cx1,cy0,cy1),(cz0,cz1,rho0,rho1)}; oc=c-o; b=oc.d; D=b^2-(oc.oc-rho)
template <int R>
__global__ void __launch_bounds__(256) k_filter_general(const float4 *tab, int npairs, int iters, unsigned *out) {
  extern __shared__ float4 s[];
  for (int i = threadIdx.x; i < 2 * npairs; i += blockDim.x) s[i] = tab[i];
  __syncthreads();
  float2 dx[R], dy[R], dz[R], ox[R], oy[R], oz[R];
  unsigned acc[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    float u = (threadIdx.x * R + r) * 1e-4f + blockIdx.x * 1e-3f;
    dx[r] = make_float2(u, u); dy[r] = make_float2(0.3f - u, 0.3f - u); dz[r] = make_float2(-0.9f, -0.9f);
    ox[r] = make_float2(-u, -u); oy[r] = make_float2(-1.f, -1.f); oz[r] = make_float2(-2.f, -2.f); acc[r] = 0xffffffffu;
  }
  unsigned slow = 0;
  for (int it = 0; it < iters; it++) {
    for (int i = 0; i < npairs; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float4 a = s[2 * (i + k)], b = s[2 * (i + k) + 1];
#pragma unroll
        for (int r = 0; r < R; r++) {
          float2 x = __fadd2_rn(make_float2(a.x, a.y), ox[r]);
          float2 y = __fadd2_rn(make_float2(a.z, a.w), oy[r]);
          float2 z = __fadd2_rn(make_float2(b.x, b.y), oz[r]);
          float2 t = __fmul2_rn(x, dx[r]);
          t = __ffma2_rn(y, dy[r], t);
          t = __ffma2_rn(z, dz[r], t);
          float2 q = __ffma2_rn(x, x, make_float2(-b.z, -b.w));
          q = __ffma2_rn(y, y, q);
          q = __ffma2_rn(z, z, q);
          float2 D = __ffma2_rn(t, t, make_float2(-q.x, -q.y));
          acc[r] &= __float_as_uint(D.x) & __float_as_uint(D.y);
        }
      }
      unsigned m = acc[0];
#pragma unroll
      for (int r = 1; r < R; r++) m &= acc[r];
      if ((int)m >= 0) {
        slow++;
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0xffffffffu;
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) dx[r].x += 1e-7f;
  }
  unsigned o = slow;
#pragma unroll
  for (int r = 0; r < R; r++) o += acc[r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = o;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  void *buf; CK(cudaMalloc(&buf, (size_t)sms * 8 * 1024 * 16));
  unsigned long long *clk; CK(cudaMalloc(&clk, 16));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_rate_khz\": %d", p.name, sms, p.major, p.minor, p.clockRate);

  // --- issue peaks: 4 CTAs x 256 threads per SM, ILP 8
  const int iters = 4096;
  const int grid = sms * 8, block = 256;
  {
    float ms = time_ms([&] { k_ffma<8><<<grid, block>>>((float *)buf, iters, 1.0001f, 0.5f, clk); });
    unsigned long long h[2]; CK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
    double flops = 2.0 * grid * block * (double)iters * 16 * 8;
    printf(", \"sm_clock_mhz_under_ffma\": %.0f", (double)h[0] / (double)h[1] * 1e3);
    printf(", \"ffma_tflops\": %.2f", flops / ms * 1e-9);
    printf(", \"ffma_per_clk_per_sm\": %.1f", flops / 2 / ms * 1e-3 / ((double)h[0] / (double)h[1] * 1e9) / sms);
  }
  {
    float ms = time_ms([&] { k_ffma2<8><<<grid, block>>>((float2 *)buf, iters, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f)); });
    double flops = 4.0 * grid * block * (double)iters * 16 * 8;
    printf(", \"ffma2_tflops\": %.2f", flops / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { k_dfma<8><<<grid, block>>>((double *)buf, iters / 4, 1.0001, 0.5); });
    double flops = 2.0 * grid * block * (double)(iters / 4) * 16 * 8;
    printf(", \"dfma_tflops\": %.2f", flops / ms * 1e-9);
  }

  // --- filter loops: 154 spheres (77 pairs -> padded to 80 / 160), 2 CTAs x 256 thr per SM resident
  const int n = 160, npairs = 80;
  std::vector<float4> tab(2 * npairs);
  for (int i = 0; i < n; i++) {  // far-away small spheres: D < 0 nearly always
    float4 v = make_float4(5.f + 0.1f * i, 3.f - 0.05f * i, -20.f - 0.2f * i, -(500.f + i));
    tab[i] = v;
  }
  float4 *dtab; CK(cudaMalloc(&dtab, tab.size() * sizeof(float4)));
  CK(cudaMemcpy(dtab, tab.data(), tab.size() * sizeof(float4), cudaMemcpyHostToDevice));
  std::vector<float4> ptab(2 * npairs);
  for (int q = 0; q < npairs; q++) {
    float4 a = tab[2 * q], b = tab[2 * q + 1];
    ptab[2 * q] = make_float4(a.x, b.x, a.y, b.y);
    ptab[2 * q + 1] = make_float4(a.z, b.z, a.w, b.w);
  }
  float4 *dptab; CK(cudaMalloc(&dptab, ptab.size() * sizeof(float4)));
  CK(cudaMemcpy(dptab, ptab.data(), ptab.size() * sizeof(float4), cudaMemcpyHostToDevice));
  const int fit = 400, fgrid = sms * 8;
  auto report = [&](const char *name, int R, float ms) {
    double tests = (double)fgrid * 256 * R * n * (double)fit;
    printf(", \"%s_R%d_Gtests_s\": %.1f", name, R, tests / ms * 1e-6);
  };
#define RUN_S(R) report("scalar", R, time_ms([&] { k_filter_scalar<R><<<fgrid, 256, n * 16>>>(dtab, n, fit, (unsigned *)buf); }))
#define RUN_P(R) report("pairs", R, time_ms([&] { k_filter_pairs<R><<<fgrid, 256, n * 16>>>(dptab, npairs, fit, (unsigned *)buf); }))
#define RUN_G(R) report("general", R, time_ms([&] { k_filter_general<R><<<fgrid, 256, n * 16>>>(dptab, npairs, fit, (unsigned *)buf); }))
  RUN_S(1); RUN_S(2); RUN_S(4);
  RUN_P(1); RUN_P(2); RUN_P(4);
  RUN_G(1); RUN_G(2);
  printf("}\n");
  return 0;
}
