// Input preparation kernels: permutation-symmetry verification of the ERI tensor and the
// symmetrised / padded / permuted copy of the 2-RDM consumed by k_gamma_contract.
#pragma once
#include "oo_common.cuh"

namespace oo {

// max over all elements of |g[pqrs]-g[qpsr]|, |g[pqrs]-g[rspq]|, |g[pqrs]-g[srqp]| and of |g|.
// out[0] = asymmetry, out[1] = max |g| (as bit patterns of non-negative doubles -> atomicMax on
// unsigned long long is order preserving).
__global__ void k_v4_symmetry(const double* __restrict__ g, int M, unsigned long long* out) {
  const size_t M2 = (size_t)M * M, M3 = M2 * M, total = M3 * M;
  double asym = 0.0, amax = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(idx % M);
    const int r = (int)((idx / M) % M);
    const int q = (int)((idx / M2) % M);
    const int p = (int)(idx / M3);
    const double v = g[idx];
    amax = fmax(amax, fabs(v));
    asym = fmax(asym, fabs(v - g[(size_t)q * M3 + (size_t)p * M2 + (size_t)s * M + r]));
    asym = fmax(asym, fabs(v - g[(size_t)r * M3 + (size_t)s * M2 + (size_t)p * M + q]));
    asym = fmax(asym, fabs(v - g[(size_t)s * M3 + (size_t)r * M2 + (size_t)q * M + p]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    asym = fmax(asym, __shfl_xor_sync(0xffffffffu, asym, o));
    amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out + 0, (unsigned long long)__double_as_longlong(asym));
    atomicMax(out + 1, (unsigned long long)__double_as_longlong(amax));
  }
}

// Gp[a][j][l*Np+k] = (1/4)(G[a,j,k,l] + G[j,a,l,k] + G[k,l,a,j] + G[l,k,j,a])  (V4 average),
// zero in the padding (j, k or l >= N).  grid N*Np, block Np*Np threads (looped).
__global__ void k_prepare_gamma(const double* __restrict__ G, double* __restrict__ Gp, int N, int Np,
                                int symmetrise) {
  const int a = blockIdx.x / Np, j = blockIdx.x % Np;
  const int Np2 = Np * Np;
  const size_t N2 = (size_t)N * N, N3 = N2 * N;
  for (int e = threadIdx.x; e < Np2; e += blockDim.x) {
    const int l = e / Np, k = e - l * Np;
    double v = 0.0;
    if (j < N && k < N && l < N) {
      v = G[a * N3 + j * N2 + (size_t)k * N + l];
      if (symmetrise) {
        v += G[j * N3 + a * N2 + (size_t)l * N + k];
        v += G[k * N3 + l * N2 + (size_t)a * N + j];
        v += G[l * N3 + k * N2 + (size_t)j * N + a];
        v *= 0.25;
      }
    }
    Gp[((size_t)a * Np + j) * Np2 + e] = v;
  }
}

}  // namespace oo
