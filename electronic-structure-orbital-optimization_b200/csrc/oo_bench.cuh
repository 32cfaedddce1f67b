// Micro-benchmarks that measure the roofline denominators on the device the library runs on:
// the FP64 tensor pipe (DMMA.8x8x4), the FP64 FMA pipe and a plain streaming read.
#pragma once
#include "oo_common.cuh"

namespace oo {

// Each warp runs `iters` x 16 independent DMMA chains from registers.  flops = 512 per DMMA.
__global__ void __launch_bounds__(256) k_peak_dmma(double* out, int iters) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  if (s == 12345.678) out[0] = s;  // keep the chain alive without real traffic
}

__global__ void __launch_bounds__(256) k_peak_dfma(double* out, int iters) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 1e-3 * i;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(512) k_stream_read(const double2* __restrict__ in, size_t n2,
                                                     double* out) {
  double s = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    const double2 a = __ldg(in + i), b = __ldg(in + i + stride), c = __ldg(in + i + 2 * stride),
                  d = __ldg(in + i + 3 * stride);
    s += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
  }
  for (; i < n2; i += stride) {
    const double2 a = __ldg(in + i);
    s += a.x + a.y;
  }
  if (s == 12345.678) out[0] = s;
}

}  // namespace oo
