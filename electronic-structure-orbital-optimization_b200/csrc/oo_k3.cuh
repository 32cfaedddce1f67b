// K3 — Stiefel retraction, Barzilai-Borwein step and the stopping-rule state machine, fused into a
// single one-CTA kernel so that an optimiser iteration needs no host round trip.
//
// Restates (reference: electronic_structure_algorithms/orbital_optimization/
//           partial_unitary_projection_optimizer.py)
//   orth(V) = V Q diag(L)^(-1/2) Q^T, (L,Q) = eigh(V^T V)                     :70-83
//   BB step:  odd k   alpha = <dU,dU> / |<dU,dG>|                              :143-148
//             even k  alpha = |<dU,dG>| / <dG,dG>   (k != 0)                   :150-155
//             U_{k+1} = orth(U_k - alpha G_k)                                  :157
//   driver:   three hand-unrolled iterations then `while S[0] > tol and k <= maxiter`,
//             S_k = (1-d)|dE| + d S_{k-1} with the one-step lag of the loop body   :176-350
// (V^T V)^(-1/2) (N <= 32) comes from a coupled Newton-Schulz iteration in shared memory; if that
// does not converge (very ill-conditioned Gram matrix) a cyclic two-sided Jacobi eigensolver
// (round-robin pairing, N/2 rotations in parallel) takes over.
#pragma once
#include "oo_common.cuh"

namespace oo {

constexpr int K3_THREADS = 1024;  // one CTA, latency-bound: short per-thread loops
constexpr int K3_NMAX = 32;

struct OptState {
  double alpha;      // current BB step size (reference: self._BBstepsize)
  double P4[3];      // [f(U_k), f(U_{k-1}), f(U_{k-2})]
  double S[2];       // [S_k, S_{k-1}]
  double tol;        // stopping_tolerance
  double decay;      // decay_factor
  double E_final;    // reference return value P4_array[0]
  int k;             // iteration_number
  int maxiter;
  int done;          // stop flag (kernels become no-ops when set)
  int k_final;       // iteration_number at loop exit
  int nan_flag;      // set when a non-finite value is met
  int ns_iters;      // telemetry: Newton-Schulz iterations summed over all retractions
  int jacobi_calls;  // telemetry: retractions that fell back to the Jacobi eigensolver
  int pad;
};

// Symmetric eigen-decomposition A = Q diag(w) Q^T of an n x n matrix held in shared memory.
// On exit A holds the (almost) diagonal matrix and Q the eigenvectors as columns.
// Parallel cyclic Jacobi, all threads of the CTA must call.  lda = K3_NMAX + 1.
__device__ inline void jacobi_eigh_smem(double* A, double* Q, double* cs, int n, int* sflag) {
  constexpr int LD = K3_NMAX + 1;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int ne = (n + 1) & ~1;  // players in the round-robin tournament (dummy if n is odd)
  for (int idx = tid; idx < n * n; idx += nth) {
    const int i = idx / n, j = idx - i * n;
    Q[i * LD + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int sweep = 0; sweep < 30; ++sweep) {
    // convergence: off-diagonal Frobenius norm relative to the diagonal
    if (tid == 0) {
      double off = 0.0, dia = 0.0;
      for (int i = 0; i < n; ++i) {
        dia += A[i * LD + i] * A[i * LD + i];
        for (int j = i + 1; j < n; ++j) off += A[i * LD + j] * A[i * LD + j];
      }
      *sflag = (off <= 1e-31 * dia) ? 1 : 0;
    }
    __syncthreads();
    if (*sflag) break;
    for (int round = 0; round < ne - 1; ++round) {
      // pair m of this round: players (round-robin "circle" method)
      if (tid < ne / 2) {
        int a = (tid == 0) ? ne - 1 : (round + tid) % (ne - 1);
        int b = (round + ne - 1 - tid) % (ne - 1);
        int p = min(a, b), q = max(a, b);
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[p * LD + q];
          if (apq != 0.0) {
            const double app = A[p * LD + p], aqq = A[q * LD + q];
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        }
        cs[4 * tid + 0] = c;
        cs[4 * tid + 1] = s;
        cs[4 * tid + 2] = (double)p;
        cs[4 * tid + 3] = (double)q;
      }
      __syncthreads();
      // A <- J^T A  (rows p,q)
      for (int idx = tid; idx < (ne / 2) * n; idx += nth) {
        const int m = idx / n, col = idx - m * n;
        const int p = (int)cs[4 * m + 2], q = (int)cs[4 * m + 3];
        if (q < n) {
          const double c = cs[4 * m], s = cs[4 * m + 1];
          const double x = A[p * LD + col], y = A[q * LD + col];
          A[p * LD + col] = c * x - s * y;
          A[q * LD + col] = s * x + c * y;
        }
      }
      __syncthreads();
      // A <- A J (columns p,q);  Q <- Q J
      for (int idx = tid; idx < (ne / 2) * n; idx += nth) {
        const int m = idx / n, row = idx - m * n;
        const int p = (int)cs[4 * m + 2], q = (int)cs[4 * m + 3];
        if (q < n) {
          const double c = cs[4 * m], s = cs[4 * m + 1];
          const double x = A[row * LD + p], y = A[row * LD + q];
          A[row * LD + p] = c * x - s * y;
          A[row * LD + q] = s * x + c * y;
          const double u = Q[row * LD + p], v = Q[row * LD + q];
          Q[row * LD + p] = c * u - s * v;
          Q[row * LD + q] = s * u + c * v;
        }
      }
      __syncthreads();
    }
  }
}

// Z -> (A/c)^(-1/2) by the coupled Newton-Schulz iteration
//     T = Z Y,  Y <- Y (3I - T)/2,  Z <- (3I - T) Z / 2,     Y0 = A/c, Z0 = I,
// c = max row sum of |A| (>= lambda_max, so the spectrum of A/c lies in (0,1] and the iteration
// converges; quadratically once ||I - T|| < 1).  Only N x N products: ~3 us where the Jacobi
// eigensolver needs ~100 us.  Returns false if 60 iterations did not reach ||I - T||_F < 1e-8
// (then the caller falls back to the eigensolver).  All threads of the CTA must call.
template <int THREADS>
__device__ inline bool newton_schulz_invsqrt(const double* A, double* Y, double* Z, double* W,
                                             int n, double* scratch, double* inv_sqrt_c,
                                             int* iters_out) {
  constexpr int LD = K3_NMAX + 1;
  constexpr int EPT = (K3_NMAX * K3_NMAX + THREADS - 1) / THREADS;  // elements per thread
  const int tid = threadIdx.x, nth = blockDim.x, nn = n * n;
  if (n <= 8) {
    // Small active spaces (configs 1-3: N = 2, 4): the matrices have at most 64 entries, two per
    // lane of ONE warp, so the whole iteration runs warp-synchronously (__syncwarp instead of
    // five CTA-wide barriers per step; the step is pure latency at these sizes).  Same arithmetic
    // as below; the residual is summed by a fixed shuffle tree.
    if (tid < 32) {
      const int lane = tid;
      double rs = 0.0, dv = 0.0;
      if (lane < n) {
        for (int j = 0; j < n; ++j) rs += fabs(A[lane * LD + j]);
        dv = rs - A[lane * LD + lane] + fabs(A[lane * LD + lane] - 1.0);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        rs = fmax(rs, __shfl_xor_sync(0xffffffffu, rs, o));
        dv = fmax(dv, __shfl_xor_sync(0xffffffffu, dv, o));
      }
      const double cw = dv < 0.5 ? 1.0 : rs;
      const double inv_cw = 1.0 / cw;
      int idx2[2], ii[2], jj[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        idx2[k] = lane + 32 * k;
        ii[k] = idx2[k] / n;
        jj[k] = idx2[k] - ii[k] * n;
        if (idx2[k] < nn) {
          Y[ii[k] * LD + jj[k]] = A[ii[k] * LD + jj[k]] * inv_cw;
          Z[ii[k] * LD + jj[k]] = (ii[k] == jj[k]) ? 1.0 : 0.0;
        }
      }
      __syncwarp();
      bool conv = false;
      int its = 0;
      for (int it = 0; it < 60; ++it) {
        double r = 0.0;
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (idx2[k] < nn) {
            double t = 0.0;
            for (int m = 0; m < n; ++m) t = fma(Z[ii[k] * LD + m], Y[m * LD + jj[k]], t);
            const double d = t - ((ii[k] == jj[k]) ? 1.0 : 0.0);
            r = fma(d, d, r);
            W[ii[k] * LD + jj[k]] = ((ii[k] == jj[k]) ? 3.0 : 0.0) - t;
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        __syncwarp();
        double yv[2] = {0.0, 0.0}, zv[2] = {0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (idx2[k] < nn) {
            double sy = 0.0, sz = 0.0;
            for (int m = 0; m < n; ++m) {
              sy = fma(Y[ii[k] * LD + m], W[m * LD + jj[k]], sy);
              sz = fma(W[ii[k] * LD + m], Z[m * LD + jj[k]], sz);
            }
            yv[k] = 0.5 * sy;
            zv[k] = 0.5 * sz;
          }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (idx2[k] < nn) {
            Y[ii[k] * LD + jj[k]] = yv[k];
            Z[ii[k] * LD + jj[k]] = zv[k];
          }
        __syncwarp();
        its = it + 1;
        if (r < 1e-16) {
          conv = true;
          break;
        }
      }
      if (lane == 0) {
        scratch[0] = conv ? 1.0 : 0.0;
        scratch[1] = cw;
        scratch[2] = (double)its;
      }
    }
    __syncthreads();
    const bool conv_all = scratch[0] != 0.0;
    const double c_all = scratch[1];
    if (iters_out) *iters_out = (int)scratch[2];
    __syncthreads();                       // scratch is reused by the caller's next reduction
    *inv_sqrt_c = 1.0 / sqrt(c_all);
    return conv_all;
  }
  if (tid < n) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += fabs(A[tid * LD + j]);
    scratch[tid] = s;
  }
  __syncthreads();
  double c = 0.0, dev = 0.0;
  for (int i = 0; i < n; ++i) {
    c = fmax(c, scratch[i]);
    // |A - I| row sum: scratch holds sum_j |A_ij|; with A_ii > 0 the deviation is that sum minus
    // A_ii plus |A_ii - 1|
    dev = fmax(dev, scratch[i] - A[i * LD + i] + fabs(A[i * LD + i] - 1.0));
  }
  __syncthreads();
  // Inside the optimiser V = U - alpha G with U orthonormal, so A = V^T V is close to the identity:
  // when ||A - I||_inf < 1/2 the iteration converges quadratically from the unscaled matrix
  // (3-4 steps instead of 7-9 after the division by the row-sum bound, which pushes the spectrum
  // away from 1).  The general case keeps the safe scaling.
  if (dev < 0.5) c = 1.0;
  const double inv_c = 1.0 / c;
  for (int idx = tid; idx < nn; idx += nth) {
    const int i = idx / n, j = idx - i * n;
    Y[i * LD + j] = A[i * LD + j] * inv_c;
    Z[i * LD + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  bool converged = false;
  for (int it = 0; it < 60; ++it) {
    double r = 0.0;
    for (int idx = tid; idx < nn; idx += nth) {
      const int i = idx / n, j = idx - i * n;
      double s = 0.0;
#pragma unroll 4
      for (int m = 0; m < n; ++m) s = fma(Z[i * LD + m], Y[m * LD + j], s);
      const double d = s - ((i == j) ? 1.0 : 0.0);
      r = fma(d, d, r);
      W[i * LD + j] = ((i == j) ? 3.0 : 0.0) - s;
    }
    r = block_sum(r, scratch);  // syncs: W is complete afterwards
    double yv[EPT], zv[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const int idx = tid + k * nth;
      yv[k] = zv[k] = 0.0;
      if (idx < nn) {
        const int i = idx / n, j = idx - i * n;
        double sy = 0.0, sz = 0.0;
#pragma unroll 4
        for (int m = 0; m < n; ++m) {
          sy = fma(Y[i * LD + m], W[m * LD + j], sy);
          sz = fma(W[i * LD + m], Z[m * LD + j], sz);
        }
        yv[k] = 0.5 * sy;
        zv[k] = 0.5 * sz;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const int idx = tid + k * nth;
      if (idx < nn) {
        const int i = idx / n, j = idx - i * n;
        Y[i * LD + j] = yv[k];
        Z[i * LD + j] = zv[k];
      }
    }
    __syncthreads();
    if (iters_out) *iters_out = it + 1;
    if (r < 1e-16) {  // ||I - T||_F < 1e-8 before this update => ~1e-16 after it
      converged = true;
      break;
    }
  }
  *inv_sqrt_c = 1.0 / sqrt(c);
  return converged;
}

// U_out = orth(V) = V (V^T V)^(-1/2) for an M x N matrix V in global memory.  One CTA.  V and
// U_out must be distinct buffers.  smem: sA, sB1, sB2, sB3 are K3_NMAX*(K3_NMAX+1) doubles each; cs 4*K3_NMAX;
// scratch 32 doubles.
// V is given as Va - alpha * Vb (Vb may be NULL: V = Va), so that the optimiser step needs no
// intermediate copy of V in global memory (one dependent round trip less).
template <int THREADS>
__device__ inline void retract_cta(const double* Va, const double* Vb, double alpha, double* Uout,
                                   int M, int N, double* sA, double* sB1, double* sB2,
                                   double* sB3, double* cs, double* scratch, int* sflag,
                                   int* telemetry = nullptr, bool force_jacobi = false) {
  constexpr int LD = K3_NMAX + 1;
  const int tid = threadIdx.x, nth = blockDim.x;
  auto V = [&](size_t idx) { return Vb ? fma(-alpha, Vb[idx], Va[idx]) : Va[idx]; };
  // Gram matrix V^T V: the t-range is split over nth / N^2 thread groups, partials summed in
  // fixed order through shared memory (sB3 is free until the Newton-Schulz iteration starts)
  {
    const int nn = N * N;
    const int tsplit = max(1, min(nth / nn, (K3_NMAX * (K3_NMAX + 1)) / nn));
    const int chunk = (M + tsplit - 1) / tsplit;
    for (int idx = tid; idx < nn * tsplit; idx += nth) {
      const int e = idx % nn, part = idx / nn;
      const int i = e / N, j = e - i * N;
      const int tb = part * chunk, te = min(M, tb + chunk);
      double s0 = 0.0, s1 = 0.0;
      int t = tb;
#pragma unroll 4
      for (; t + 1 < te; t += 2) {
        s0 = fma(V((size_t)t * N + i), V((size_t)t * N + j), s0);
        s1 = fma(V((size_t)(t + 1) * N + i), V((size_t)(t + 1) * N + j), s1);
      }
      if (t < te) s0 = fma(V((size_t)t * N + i), V((size_t)t * N + j), s0);
      sB3[part * nn + e] = s0 + s1;
    }
    __syncthreads();
    for (int e = tid; e < nn; e += nth) {
      double s = 0.0;
      for (int part = 0; part < tsplit; ++part) s += sB3[part * nn + e];
      sA[(e / N) * LD + (e % N)] = s;
    }
    __syncthreads();
    // exact symmetry (the two triangles were summed in different orders)
    for (int e = tid; e < nn; e += nth) {
      const int i = e / N, j = e - i * N;
      if (j > i) sA[j * LD + i] = sA[i * LD + j];
    }
    __syncthreads();
  }
  double scale = 1.0;
  double* S = sB2;  // inverse square root (up to `scale`) ends up here
  int ns_it = 0;
  const bool ok = !force_jacobi &&
                  newton_schulz_invsqrt<THREADS>(sA, sB1, sB2, sB3, N, scratch, &scale, &ns_it);
  if (telemetry && tid == 0) {
    telemetry[0] += ns_it;
    if (!ok) telemetry[1] += 1;
  }
  if (!ok) {
    // robust path: S = Q diag(w^-1/2) Q^T from the Jacobi eigen-decomposition (reference:
    // torch.linalg.eigh, partial_unitary_projection_optimizer.py:80-81)
    for (int idx = tid; idx < N * N; idx += nth) {
      const int i = idx / N, j = idx - i * N;
      sB1[i * LD + j] = sA[i * LD + j];
    }
    __syncthreads();
    jacobi_eigh_smem(sB1, sB3, cs, N, sflag);
    __syncthreads();
    for (int idx = tid; idx < N * N; idx += nth) {
      const int i = idx / N, j = idx - i * N;
      double s = 0.0;
      for (int m = 0; m < N; ++m)
        s += sB3[i * LD + m] * sB3[j * LD + m] / sqrt(sB1[m * LD + m]);
      sB2[i * LD + j] = s;
    }
    scale = 1.0;
    __syncthreads();
  }
  // U = V S  (Va, Vb and Uout must not alias): nth / M threads share a row, each a subset of columns
  {
    const int per_row = max(1, nth / M);
    for (int idx = tid; idx < M * per_row; idx += nth) {
      const int t = idx / per_row, grp = idx - t * per_row;
      double v[K3_NMAX];
      for (int j = 0; j < N; ++j) v[j] = V((size_t)t * N + j) * scale;
      for (int j = grp; j < N; j += per_row) {
        double s = 0.0;
        for (int m = 0; m < N; ++m) s = fma(v[m], S[m * LD + j], s);
        Uout[(size_t)t * N + j] = s;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(K3_THREADS) k_orth(const double* V, double* Uout, int M, int N,
                                                      int force_jacobi) {
  __shared__ double sA[K3_NMAX * (K3_NMAX + 1)], sB1[K3_NMAX * (K3_NMAX + 1)],
      sB2[K3_NMAX * (K3_NMAX + 1)], sB3[K3_NMAX * (K3_NMAX + 1)], cs[4 * K3_NMAX], scratch[33];
  __shared__ int sflag;
  retract_cta<K3_THREADS>(V, nullptr, 0.0, Uout, M, N, sA, sB1, sB2, sB3, cs, scratch, &sflag,
                          nullptr, force_jacobi != 0);
}

struct StepParams {
  OptState* st;
  double* Ucur;         // U_k        -> becomes U_{k+1}
  double* Uprev;        // U_{k-1}    -> becomes U_k
  const double* gE;     // [M*N+1] gradient at U_k and f(U_k) (already all-reduced)
  double* Gprev;        // G_{k-1}    -> becomes G_k
  double* E_hist;       // E_hist[k] = f(U_k)
  int M, N;
  int hist_cap;
  int force_jacobi;     // debugging / testing: skip Newton-Schulz, use the eigensolver path
};

// Shared memory of one optimiser transition / retraction.
struct StepSmem {
  double sA[K3_NMAX * (K3_NMAX + 1)], sB1[K3_NMAX * (K3_NMAX + 1)], sB2[K3_NMAX * (K3_NMAX + 1)],
      sB3[K3_NMAX * (K3_NMAX + 1)], cs[4 * K3_NMAX], scratch[33], scratch3[3 * 33];
  int sflag;
};

// Three deterministic block-wide sums at once (fixed shuffle trees): one pair of barriers instead
// of three.  scratch3: 99 doubles.  Results valid in every thread.
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* scratch3) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  __syncthreads();
  if (lane == 0) {
    scratch3[warp] = a;
    scratch3[33 + warp] = b;
    scratch3[66 + warp] = c;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double ta = lane < nw ? scratch3[lane] : 0.0, tb = lane < nw ? scratch3[33 + lane] : 0.0,
           tc = lane < nw ? scratch3[66 + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ta += __shfl_xor_sync(0xffffffffu, ta, o);
      tb += __shfl_xor_sync(0xffffffffu, tb, o);
      tc += __shfl_xor_sync(0xffffffffu, tc, o);
    }
    if (lane == 0) {
      scratch3[32] = ta;
      scratch3[65] = tb;
      scratch3[98] = tc;
    }
  }
  __syncthreads();
  a = scratch3[32];
  b = scratch3[65];
  c = scratch3[98];
}

// The stopping-rule state machine of pupo.py:176-346 (three hand-unrolled iterations, then
// `while St_array[0] > tol and iteration_number <= maxiter`): given the state before iteration k
// and f(U_k), the histories after it and whether the loop ends here.
struct StepDecision {
  double P0, P1, P2, S0, S1;
  bool stop;
};
__device__ __forceinline__ StepDecision step_decide(const OptState& st, double fk) {
  StepDecision r{st.P4[0], st.P4[1], st.P4[2], st.S[0], st.S[1], false};
  const double d = st.decay;
  const int k = st.k;
  if (k == 0) {
    r.P2 = fk;                                          // P4_array[2] = f(U_0)
  } else if (k == 1) {
    r.P1 = fk;                                          // P4_array[1] = f(U_1)
    r.S0 = (1.0 - d) * fabs(r.P1 - r.P2) + d * r.S1;    // S_1, with S1 preset to 1.5*tol
  } else if (k == 2) {
    r.P0 = fk;                                          // P4_array[0] = f(U_2)
    r.S1 = r.S0;                                        // np.roll(St_array, 1) on a 2-vector = swap
    r.S0 = (1.0 - d) * fabs(r.P0 - r.P1) + d * r.S1;    // S_2
  } else if (!(r.S0 > st.tol && k <= st.maxiter)) {
    r.stop = true;
  } else {
    r.P2 = r.P1; r.P1 = r.P0; r.P0 = fk;                // roll, then P4_array[0] = f(U_k)
    const double s_old = r.S0;
    r.S1 = s_old;
    r.S0 = (1.0 - d) * fabs(r.P1 - r.P2) + d * r.S1;    // lagged difference, as in the reference
  }
  return r;
}

// One optimiser transition: consumes (f(U_k), G_k) and produces U_{k+1}, or raises the stop flag.
// Called by every thread of ONE CTA of THREADS threads (the stand-alone k_step kernel, or the last
// CTA of the tail kernel when the step is fused into the evaluation).
// sV (optional): M*N doubles of shared memory for V = U_k - alpha G_k.  Without it the Gram matrix
// and the final product read V through 4*M global loads per thread, which makes the one CTA
// LSU-bound (~100 us cold at M=256, N=16); from shared memory they are conflict-free LDS.
template <int THREADS>
__device__ inline void opt_step_cta(const StepParams& p, StepSmem& sm, double* sV = nullptr) {
  OptState* st = p.st;
  if (st->done) return;
  const int tid = threadIdx.x, nth = THREADS;
  const int MN = p.M * p.N;
  const OptState s0 = *st;
  const int k = s0.k;
  const double fk = p.gE[MN];
  double alpha = s0.alpha;
  const StepDecision dec = step_decide(s0, fk);
  __syncthreads();  // every thread has read the state before thread 0 rewrites it

  if (dec.stop) {
    if (tid == 0) {
      st->done = 1;
      st->k_final = k;
      st->E_final = dec.P0;                    // reference returns P4_array[0] = f(U_{k-1})
    }
    return;
  }
  if (tid == 0 && k < p.hist_cap) p.E_hist[k] = fk;

  // ---- BB step size -------------------------------------------------------
  if (k >= 1) {
    double uu = 0.0, ug = 0.0, gg = 0.0;
    for (int i = tid; i < MN; i += nth) {
      const double du = p.Ucur[i] - p.Uprev[i];
      const double dg = p.gE[i] - p.Gprev[i];
      uu = fma(du, du, uu);
      ug = fma(du, dg, ug);
      gg = fma(dg, dg, gg);
    }
    block_sum3(uu, ug, gg, sm.scratch3);
    alpha = (k & 1) ? uu / fabs(ug) : fabs(ug) / gg;
  }
  // ---- shift histories; V = U_k - alpha G_k into shared memory when there is room, else formed
  // on the fly from the shifted copies ------------------------------------------------------
  for (int i = tid; i < MN; i += nth) {
    const double u = p.Ucur[i], g = p.gE[i];
    p.Uprev[i] = u;
    p.Gprev[i] = g;
    if (sV) sV[i] = fma(-alpha, g, u);
  }
  __syncthreads();
  if (sV)
    retract_cta<THREADS>(sV, nullptr, 0.0, p.Ucur, p.M, p.N, sm.sA, sm.sB1, sm.sB2, sm.sB3, sm.cs,
                         sm.scratch, &sm.sflag, &st->ns_iters, p.force_jacobi != 0);
  else
    retract_cta<THREADS>(p.Uprev, p.Gprev, alpha, p.Ucur, p.M, p.N, sm.sA, sm.sB1, sm.sB2, sm.sB3,
                         sm.cs, sm.scratch, &sm.sflag, &st->ns_iters, p.force_jacobi != 0);
  if (tid == 0) {
    st->alpha = alpha;
    st->P4[0] = dec.P0; st->P4[1] = dec.P1; st->P4[2] = dec.P2;
    st->S[0] = dec.S0; st->S[1] = dec.S1;
    st->k = k + 1;
    if (!isfinite(alpha) || !isfinite(fk)) st->nan_flag = 1;
  }
}

// The same transition for small problems (M*N <= STEP_SMALL_MN), where the step is pure latency:
// U_k, U_{k-1}, G_{k-1} and the optimiser state were loaded into shared memory at the START of the
// tail kernel (in the shadow of the row reduction), G_k is read once, V lives in shared memory, so
// the chain of dependent global round trips shrinks from seven to one.
constexpr int STEP_SMALL_MN = 256;
struct StepSmallSmem {
  double U[STEP_SMALL_MN], Up[STEP_SMALL_MN], Gp[STEP_SMALL_MN], G[STEP_SMALL_MN];
  OptState st;
};

// issued by every thread of every CTA of the tail kernel right after its dependency wait
template <int THREADS>
__device__ __forceinline__ void opt_step_small_prefetch(const StepParams& p, StepSmallSmem& ss) {
  const int MN = p.M * p.N;
  for (int i = threadIdx.x; i < MN; i += THREADS) {
    ss.U[i] = p.Ucur[i];
    ss.Up[i] = p.Uprev[i];
    ss.Gp[i] = p.Gprev[i];
  }
  if (threadIdx.x == 0) ss.st = *p.st;
}

template <int THREADS>
__device__ inline void opt_step_small_cta(const StepParams& p, StepSmem& sm, StepSmallSmem& ss) {
  OptState* st = p.st;
  const int tid = threadIdx.x, nth = THREADS;
  const int MN = p.M * p.N;
  for (int i = tid; i < MN; i += nth) ss.G[i] = p.gE[i];     // the one dependent round trip
  const double fk = p.gE[MN];
  __syncthreads();                                           // prefetched data + G visible
  const OptState s0 = ss.st;
  if (s0.done) return;
  const int k = s0.k;
  double alpha = s0.alpha;
  const StepDecision dec = step_decide(s0, fk);
  if (dec.stop) {
    if (tid == 0) {
      st->done = 1;
      st->k_final = k;
      st->E_final = dec.P0;
    }
    return;
  }
  if (tid == 0 && k < p.hist_cap) p.E_hist[k] = fk;
  if (k >= 1) {
    double uu = 0.0, ug = 0.0, gg = 0.0;
    for (int i = tid; i < MN; i += nth) {
      const double du = ss.U[i] - ss.Up[i];
      const double dg = ss.G[i] - ss.Gp[i];
      uu = fma(du, du, uu);
      ug = fma(du, dg, ug);
      gg = fma(dg, dg, gg);
    }
    block_sum3(uu, ug, gg, sm.scratch3);
    alpha = (k & 1) ? uu / fabs(ug) : fabs(ug) / gg;
  }
  // histories to global memory (nobody waits for these stores); V = U_k - alpha G_k over U_{k-1}
  for (int i = tid; i < MN; i += nth) {
    const double u = ss.U[i], g = ss.G[i];
    p.Uprev[i] = u;
    p.Gprev[i] = g;
    ss.Up[i] = fma(-alpha, g, u);
  }
  __syncthreads();
  retract_cta<THREADS>(ss.Up, nullptr, 0.0, p.Ucur, p.M, p.N, sm.sA, sm.sB1, sm.sB2, sm.sB3, sm.cs,
                       sm.scratch, &sm.sflag, &st->ns_iters, p.force_jacobi != 0);
  if (tid == 0) {
    st->alpha = alpha;
    st->P4[0] = dec.P0; st->P4[1] = dec.P1; st->P4[2] = dec.P2;
    st->S[0] = dec.S0; st->S[1] = dec.S1;
    st->k = k + 1;
    if (!isfinite(alpha) || !isfinite(fk)) st->nan_flag = 1;
  }
}

// v_in_smem: the launch carries M*N doubles of dynamic shared memory for V (see opt_step_cta).
__global__ void __launch_bounds__(K3_THREADS) k_step(const StepParams p, int v_in_smem) {
  __shared__ StepSmem sm;
  extern __shared__ double k_step_v[];
  opt_step_cta<K3_THREADS>(p, sm, v_in_smem ? k_step_v : nullptr);
}

// Early stop requested by the host (a callback raised): behaves like the loop exit of the
// reference at the current iteration (return value P4_array[0]).
__global__ void k_force_stop(OptState* st) {
  if (st->done) return;
  st->done = 1;
  st->k_final = st->k;
  st->E_final = st->P4[0];
}

// Standalone BB update (compute_updated_partial_unitary, pupo.py:129-159) for API parity.
struct BBParams {
  const double* Ucur; const double* Uprev; const double* Gcur; const double* Gprev;
  double* Unew; double* alpha_io; int iteration; int M, N; int force_jacobi;
};
__global__ void __launch_bounds__(K3_THREADS) k_bb_update(const BBParams p) {
  __shared__ double sA[K3_NMAX * (K3_NMAX + 1)], sB1[K3_NMAX * (K3_NMAX + 1)],
      sB2[K3_NMAX * (K3_NMAX + 1)], sB3[K3_NMAX * (K3_NMAX + 1)], cs[4 * K3_NMAX], scratch[33];
  __shared__ int sflag;
  const int tid = threadIdx.x, nth = blockDim.x, MN = p.M * p.N;
  double alpha = *p.alpha_io;
  __syncthreads();
  if (p.iteration >= 1) {
    double uu = 0.0, ug = 0.0, gg = 0.0;
    for (int i = tid; i < MN; i += nth) {
      const double du = p.Ucur[i] - p.Uprev[i];
      const double dg = p.Gcur[i] - p.Gprev[i];
      uu = fma(du, du, uu);
      ug = fma(du, dg, ug);
      gg = fma(dg, dg, gg);
    }
    uu = block_sum(uu, scratch);
    ug = block_sum(ug, scratch);
    gg = block_sum(gg, scratch);
    alpha = (p.iteration & 1) ? uu / fabs(ug) : fabs(ug) / gg;
  }
  retract_cta<K3_THREADS>(p.Ucur, p.Gcur, alpha, p.Unew, p.M, p.N, sA, sB1, sB2, sB3, cs, scratch,
                          &sflag, nullptr, p.force_jacobi != 0);
  if (tid == 0) *p.alpha_io = alpha;
}

}  // namespace oo
