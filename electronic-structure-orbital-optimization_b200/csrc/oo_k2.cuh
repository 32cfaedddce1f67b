// K2 family — the cheap tail of the energy/gradient evaluation (HBM/L2-bound, no tensor cores):
//   k_qcontract      T3[t][j][e]  = sum_q U[q][j] * Y[t,q][e]                (third index contraction)
//   k_gamma_contract A[t][a]      = sum_{j,e} T3[t][j][e] * Gp[a][j][e]      (2-RDM contraction)
//   k_ud             UD = U*D, UDt = U*D^T                                    (1-RDM, tiny)
//   k_finalize       dE/dU rows of this GPU's shard, partial energy, fixed-order reductions
//   k_rotate_g       g'[i][j][k][l] = sum_t U[t][i] * T3[t][j][k][l]          (rotated Hamiltonian)
// Together with K1 they restate, in the spatial-orbital picture and with an analytic gradient,
//   base_opt_orb_solver.py:554-563 (energy) and
//   partial_unitary_projection_optimizer.py:85-103 (autograd gradient) of the reference.
// Layout: e = l*Np + k indexes the padded (Np = 8*NT) N x N tile exactly as K1 stores it.
#pragma once
#include "oo_common.cuh"

namespace oo {

constexpr int QC_ECHUNK = 64;  // e-values per CTA in k_qcontract
constexpr int QC_QGROUPS = 4;  // q-range split inside the CTA

// grid (Mloc, ceil(Np^2/64)), block 256, dynamic smem: M*Np doubles (U padded) + 4*Np*64 doubles
template <int NT>
__global__ void __launch_bounds__(QC_ECHUNK* QC_QGROUPS)
k_qcontract(const double* __restrict__ Y, const double* __restrict__ U, double* __restrict__ T3,
            int M, int N, const int* done_flag) {
  constexpr int Np = NT * 8, Np2 = Np * Np;
  if (done_flag != nullptr && *done_flag != 0) return;
  extern __shared__ double qc_smem[];
  double* Us = qc_smem;                 // [M][Np]
  double* red = qc_smem + (size_t)M * Np;  // [QGROUPS][Np][ECHUNK]
  const int tid = threadIdx.x;
  for (int idx = tid; idx < M * Np; idx += blockDim.x) {
    const int q = idx / Np, j = idx - q * Np;
    Us[idx] = (j < N) ? __ldg(U + (size_t)q * N + j) : 0.0;
  }
  __syncthreads();
  const int t = blockIdx.x;
  const int el = tid & (QC_ECHUNK - 1), qg = tid / QC_ECHUNK;
  const int e = blockIdx.y * QC_ECHUNK + el;
  const bool valid = e < Np2;
  double acc[Np];
#pragma unroll
  for (int j = 0; j < Np; ++j) acc[j] = 0.0;
  const int qper = (M + QC_QGROUPS - 1) / QC_QGROUPS;
  const int q0 = qg * qper, q1 = min(M, q0 + qper);
  const double* yp = Y + ((size_t)t * M) * Np2 + (valid ? e : 0);
#pragma unroll 4
  for (int q = q0; q < q1; ++q) {
    const double y = valid ? __ldg(yp + (size_t)q * Np2) : 0.0;
    const double2* u2 = reinterpret_cast<const double2*>(Us + q * Np);
#pragma unroll
    for (int j = 0; j < Np / 2; ++j) {
      const double2 u = u2[j];
      acc[2 * j] = fma(u.x, y, acc[2 * j]);
      acc[2 * j + 1] = fma(u.y, y, acc[2 * j + 1]);
    }
  }
#pragma unroll
  for (int j = 0; j < Np; ++j) red[(qg * Np + j) * QC_ECHUNK + el] = acc[j];
  __syncthreads();
  // fixed-order sum over the q-groups; thread (qg, el) finishes planes j = qg, qg+4, ...
  if (valid) {
    for (int j = qg; j < Np; j += QC_QGROUPS) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < QC_QGROUPS; ++w) s += red[(w * Np + j) * QC_ECHUNK + el];
      T3[((size_t)t * Np + j) * Np2 + e] = s;
    }
  }
}

constexpr int GC_AGROUP = 4;  // a-values per CTA in k_gamma_contract

// grid (Mloc, ceil(N/4)), block 256.  L = Np^3.  A is [Mloc][N].
__global__ void __launch_bounds__(256)
k_gamma_contract(const double* __restrict__ T3, const double* __restrict__ Gp,
                 double* __restrict__ A, int N, int L, const int* done_flag) {
  if (done_flag != nullptr && *done_flag != 0) return;
  __shared__ double scratch[32];
  const int t = blockIdx.x, a0 = blockIdx.y * GC_AGROUP;
  const double* tp = T3 + (size_t)t * L;
  double acc[GC_AGROUP] = {0.0, 0.0, 0.0, 0.0};
  const double* gp[GC_AGROUP];
#pragma unroll
  for (int i = 0; i < GC_AGROUP; ++i) gp[i] = Gp + (size_t)min(a0 + i, N - 1) * L;
  for (int idx = threadIdx.x * 2; idx < L; idx += blockDim.x * 2) {
    const double2 tv = *reinterpret_cast<const double2*>(tp + idx);
#pragma unroll
    for (int i = 0; i < GC_AGROUP; ++i) {
      const double2 gv = __ldg(reinterpret_cast<const double2*>(gp[i] + idx));
      acc[i] = fma(tv.x, gv.x, acc[i]);
      acc[i] = fma(tv.y, gv.y, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < GC_AGROUP; ++i) {
    const double s = block_sum(acc[i], scratch);
    if (threadIdx.x == 0 && a0 + i < N) A[(size_t)t * N + a0 + i] = s;
  }
}

// UD[q][a] = sum_j U[q][j] D[a][j]... see below.  grid M, block 32*ceil(N/32) (>= N threads).
//   UDt[q][a] = sum_j U[q][j] * D[a][j]   (= U D^T)
//   UD [q][a] = sum_j U[q][j] * D[j][a]   (= U D)
__global__ void k_ud(const double* __restrict__ U, const double* __restrict__ D,
                     double* __restrict__ UD, double* __restrict__ UDt, int N,
                     const int* done_flag) {
  if (done_flag != nullptr && *done_flag != 0) return;
  const int q = blockIdx.x, a = threadIdx.x;
  if (a >= N) return;
  double s0 = 0.0, s1 = 0.0;
  for (int j = 0; j < N; ++j) {
    const double u = U[(size_t)q * N + j];
    s0 = fma(u, D[j * N + a], s0);
    s1 = fma(u, D[a * N + j], s1);
  }
  UD[(size_t)q * N + a] = s0;
  UDt[(size_t)q * N + a] = s1;
}

struct FinalizeParams {
  const double* h;    // [M][M]
  const double* U;    // [M][N]
  const double* UD;   // [M][N]
  const double* UDt;  // [M][N]
  const double* A;    // [Mloc][N]
  double* out;        // [M*N + 1]: gradient rows (only this shard's rows are written) + energy
  double* rowE;       // [Mloc]
  unsigned int* counter;
  const int* done_flag;
  int M, N, t0, Mloc;
  double two_body_grad_factor;  // 4 for the one-pass (V4-symmetric) gradient
};

// grid Mloc, block 128.  Row t = t0 + blockIdx.x of
//   dE/dU = h (U D^T) + h^T (U D) + 4 A,      E_partial = sum_{t in shard} U[t,:].(h U D^T + A)[t,:]
__global__ void __launch_bounds__(128) k_finalize(const FinalizeParams p) {
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  __shared__ double r1[128], r2[128];
  __shared__ bool is_last;
  const int tl = blockIdx.x, t = p.t0 + tl, N = p.N, M = p.M;
  const int nparts = 128 / N > 0 ? 128 / N : 1;  // N <= 32 -> at least 4 parts
  const int a = threadIdx.x % N, part = threadIdx.x / N;
  double s1 = 0.0, s2 = 0.0;
  if (part < nparts) {
    for (int q = part; q < M; q += nparts) {
      s1 = fma(p.h[(size_t)t * M + q], p.UDt[(size_t)q * N + a], s1);
      s2 = fma(p.h[(size_t)q * M + t], p.UD[(size_t)q * N + a], s2);
    }
  }
  r1[threadIdx.x] = s1;
  r2[threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.x < N) {
    double b1 = 0.0, b2 = 0.0;
    for (int w = 0; w < nparts; ++w) {
      b1 += r1[w * N + threadIdx.x];
      b2 += r2[w * N + threadIdx.x];
    }
    const double av = p.A[(size_t)tl * N + threadIdx.x];
    p.out[(size_t)t * N + threadIdx.x] = p.two_body_grad_factor * av + b1 + b2;
    r1[threadIdx.x] = p.U[(size_t)t * N + threadIdx.x] * (av + b1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0.0;
    for (int i = 0; i < N; ++i) e += r1[i];
    p.rowE[tl] = e;
    __threadfence();
    const unsigned int prev = atomicAdd(p.counter, 1u);
    is_last = (prev == (unsigned int)(gridDim.x - 1));
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double e = 0.0;
    for (int i = 0; i < p.Mloc; ++i) e += ((volatile double*)p.rowE)[i];
    p.out[(size_t)M * N] = e;
    *p.counter = 0u;
  }
}

// g'[i][j][k][l] (N^4, unpadded, physicist order like the input) = sum_{t in shard} U[t][i]*T3[t][j][l*Np+k]
// grid (N*N) over (i,j), block Np2 threads over e. Partial over the shard when sharded.
__global__ void k_rotate_g(const double* __restrict__ T3, const double* __restrict__ U,
                           double* __restrict__ gout, int N, int Np, int t0, int Mloc) {
  const int i = blockIdx.x / N, j = blockIdx.x % N;
  const int Np2 = Np * Np;
  for (int e = threadIdx.x; e < Np2; e += blockDim.x) {
    const int l = e / Np, k = e - l * Np;
    if (l >= N || k >= N) continue;
    double s = 0.0;
    for (int tl = 0; tl < Mloc; ++tl)
      s = fma(U[(size_t)(t0 + tl) * N + i], T3[((size_t)tl * Np + j) * Np2 + e], s);
    gout[(((size_t)i * N + j) * N + k) * N + l] = s;
  }
}

// h'[i][j] = sum_{p in shard, q} U[p][i] h[p][q] U[q][j].  grid N*N, block 128.
__global__ void k_rotate_h(const double* __restrict__ h, const double* __restrict__ U,
                           double* __restrict__ hout, int M, int N, int t0, int Mloc) {
  __shared__ double scratch[32];
  const int i = blockIdx.x / N, j = blockIdx.x % N;
  double s = 0.0;
  for (int idx = threadIdx.x; idx < Mloc * M; idx += blockDim.x) {
    const int pl = idx / M, q = idx - pl * M;
    const int p = t0 + pl;
    s = fma(U[(size_t)p * N + i] * h[(size_t)p * M + q], U[(size_t)q * N + j], s);
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) hout[i * N + j] = s;
}

}  // namespace oo
