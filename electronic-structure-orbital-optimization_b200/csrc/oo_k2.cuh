// K2 family — the cheap tail of the energy/gradient evaluation (HBM/L2-bound, no tensor cores):
//   k_qcontract      T3[x][j][e]  = sum_q U[q][j] * Y[x,q][e]                (third index contraction)
//   k_gamma_contract A[x][a]      = sum_{j,e} T3[x][j][e] * Gp[a][j][e]      (2-RDM contraction)
//   k_ud             UD = U*D, UDt = U*D^T                                    (1-RDM, tiny)
//   k_finalize       dE/dU rows, partial energy, fixed-order reductions
//   k_rotate_g       g'[i][j][k][l] = sum_x U[x][i] * T3[x][j][k][l]          (rotated Hamiltonian)
// Together with K1 they restate, in the spatial-orbital picture and with an analytic gradient,
//   base_opt_orb_solver.py:554-563 (energy) and
//   partial_unitary_projection_optimizer.py:85-103 (autograd gradient) of the reference.
// Layout: e = l*Np + k indexes the padded (Np = 8*NT) N x N tile exactly as K1 stores it.
//
// Pair-symmetric mode.  For a V4-symmetric tensor g[t,q,r,s] = g[q,t,s,r], hence
// Y[q,t](k,l) = Y[t,q](l,k): K1 streams only one slab of every pair {(t,q),(q,t)} (checkerboard
// choice, so every row keeps ~M/2 slabs and shards stay balanced) and the q-contraction uses each
// computed tile twice, once as is for row t and once transposed for row q.  Row q may belong to
// another GPU: every GPU then produces partial T3/A/gradient rows for ALL x in [0,M), and the
// (M*N+1)-double all-reduce that exists anyway completes them.
#pragma once
#include "oo_common.cuh"

namespace oo {

// Is slab (t,q) the one of its pair that gets streamed?
__host__ __device__ inline bool pair_selected(int t, int q) {
  if (t == q) return true;
  return (((t + q) & 1) == 0) ? (t < q) : (t > q);
}

constexpr int QC_ECHUNK = 64;  // e-values per CTA in k_qcontract
constexpr int QC_QGROUPS = 4;  // term-range split inside the CTA

struct QCParams {
  const double* Y;       // [nslab][Np*Np]
  const double* U;       // [M][N]
  double* T3;            // [nrows][Np][Np*Np]
  const int* idxmap;     // pair-symmetric mode: [mloc][M] slab index or -1; NULL = dense (tl*M+q)
  const int* done_flag;
  int M, N, t0, mloc;
  int row0, nrows;       // rows x produced: dense [t0, t0+mloc); pair-symmetric [0, M)
};

// grid (nrows, ceil(Np^2/64)), block 256, dynamic smem: M*Np doubles (U padded) + 4*Np*64 doubles
template <int NT>
__global__ void __launch_bounds__(QC_ECHUNK* QC_QGROUPS) k_qcontract(const QCParams p) {
  constexpr int Np = NT * 8, Np2 = Np * Np;
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  extern __shared__ double qc_smem[];
  double* Us = qc_smem;                          // [M][Np]
  double* red = qc_smem + (size_t)p.M * Np;      // [QGROUPS][Np][ECHUNK]
  const int tid = threadIdx.x, M = p.M;
  for (int idx = tid; idx < M * Np; idx += blockDim.x) {
    const int q = idx / Np, j = idx - q * Np;
    Us[idx] = (j < p.N) ? __ldg(p.U + (size_t)q * p.N + j) : 0.0;
  }
  __syncthreads();
  const int x = p.row0 + blockIdx.x;
  const int el = tid & (QC_ECHUNK - 1), qg = tid / QC_ECHUNK;
  const int e = blockIdx.y * QC_ECHUNK + el;
  const bool valid = e < Np2;
  const int ee = valid ? e : 0;
  const int eT = (ee % Np) * Np + ee / Np;       // transposed position inside a tile
  double acc[Np];
#pragma unroll
  for (int j = 0; j < Np; ++j) acc[j] = 0.0;
  const bool mine = (x >= p.t0) && (x < p.t0 + p.mloc);
  // term list: n in [0,M): slab (x,q=n) of my own row; n in [M, M+mloc): slab (t=t0+n-M, x)
  // transposed (pair-symmetric mode only)
  const int nterms = p.idxmap ? M + p.mloc : M;
  const int per = (nterms + QC_QGROUPS - 1) / QC_QGROUPS;
  const int n0 = qg * per, n1 = min(nterms, n0 + per);
#pragma unroll 2
  for (int n = n0; n < n1; ++n) {
    int slab, urow, epos;
    if (n < M) {
      if (!mine) continue;
      slab = p.idxmap ? __ldg(p.idxmap + (size_t)(x - p.t0) * M + n) : (x - p.t0) * M + n;
      urow = n;
      epos = ee;
    } else {
      const int tl = n - M;
      if (p.t0 + tl == x) continue;              // the diagonal slab is already in the first list
      slab = __ldg(p.idxmap + (size_t)tl * M + x);
      urow = p.t0 + tl;
      epos = eT;
    }
    if (slab < 0) continue;
    const double y = valid ? __ldg(p.Y + (size_t)slab * Np2 + epos) : 0.0;
    const double2* u2 = reinterpret_cast<const double2*>(Us + urow * Np);
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
  // fixed-order sum over the groups; thread (qg, el) finishes planes j = qg, qg+4, ...
  if (valid) {
    for (int j = qg; j < Np; j += QC_QGROUPS) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < QC_QGROUPS; ++w) s += red[(w * Np + j) * QC_ECHUNK + el];
      p.T3[((size_t)blockIdx.x * Np + j) * Np2 + e] = s;
    }
  }
}

constexpr int GC_AGROUP = 4;  // a-values per CTA in k_gamma_contract

// grid (nrows, ceil(N/4)), block 256.  L = Np^3.  A is [nrows][N].
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

//   UDt[q][a] = sum_j U[q][j] * D[a][j]   (= U D^T)
//   UD [q][a] = sum_j U[q][j] * D[j][a]   (= U D)
// grid M, block 32 (N <= 32).
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
  const double* A;    // [nrows][N]
  double* out;        // [M*N + 1]: gradient rows + energy (rows outside [row0,row0+nrows) untouched)
  double* rowE;       // [nrows]
  unsigned int* counter;
  const int* done_flag;
  int M, N, t0, mloc;  // shard: the one-body terms are added for rows in [t0, t0+mloc) only
  int row0, nrows;
  double two_body_grad_factor;  // 4 for the one-pass (V4-symmetric) gradient
};

// grid nrows, block 128.  Row x = row0 + blockIdx.x of
//   dE/dU = 4 A + [x in shard] (h (U D^T) + h^T (U D)),   E_partial = sum_x U[x,:].(A + [..] h U D^T)[x,:]
__global__ void __launch_bounds__(128) k_finalize(const FinalizeParams p) {
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  __shared__ double r1[128], r2[128];
  __shared__ bool is_last;
  const int xl = blockIdx.x, x = p.row0 + xl, N = p.N, M = p.M;
  const bool mine = (x >= p.t0) && (x < p.t0 + p.mloc);
  const int nparts = 128 / N > 0 ? 128 / N : 1;  // N <= 32 -> at least 4 parts
  const int a = threadIdx.x % N, part = threadIdx.x / N;
  double s1 = 0.0, s2 = 0.0;
  if (mine && part < nparts) {
    for (int q = part; q < M; q += nparts) {
      s1 = fma(p.h[(size_t)x * M + q], p.UDt[(size_t)q * N + a], s1);
      s2 = fma(p.h[(size_t)q * M + x], p.UD[(size_t)q * N + a], s2);
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
    const double av = p.A[(size_t)xl * N + threadIdx.x];
    p.out[(size_t)x * N + threadIdx.x] = p.two_body_grad_factor * av + b1 + b2;
    r1[threadIdx.x] = p.U[(size_t)x * N + threadIdx.x] * (av + b1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0.0;
    for (int i = 0; i < N; ++i) e += r1[i];
    p.rowE[xl] = e;
    __threadfence();
    const unsigned int prev = atomicAdd(p.counter, 1u);
    is_last = (prev == (unsigned int)(gridDim.x - 1));
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double e = 0.0;
    for (int i = 0; i < p.nrows; ++i) e += ((volatile double*)p.rowE)[i];
    p.out[(size_t)M * N] = e;
    *p.counter = 0u;
  }
}

// g'[i][j][k][l] (N^4, unpadded, physicist order like the input)
//   = sum_{x in [row0,row0+nrows)} U[x][i] * T3[x][j][l*Np+k]
// grid (N*N) over (i,j), block 256 threads over e.  Partial over the GPU's slabs when sharded.
__global__ void k_rotate_g(const double* __restrict__ T3, const double* __restrict__ U,
                           double* __restrict__ gout, int N, int Np, int row0, int nrows) {
  const int i = blockIdx.x / N, j = blockIdx.x % N;
  const int Np2 = Np * Np;
  for (int e = threadIdx.x; e < Np2; e += blockDim.x) {
    const int l = e / Np, k = e - l * Np;
    if (l >= N || k >= N) continue;
    double s = 0.0;
    for (int xl = 0; xl < nrows; ++xl)
      s = fma(U[(size_t)(row0 + xl) * N + i], T3[((size_t)xl * Np + j) * Np2 + e], s);
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
