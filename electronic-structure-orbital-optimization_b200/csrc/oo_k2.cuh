// K2 family — everything of an evaluation around K1 (HBM/L2-bound, no tensor cores):
//   k_prepare_q      Q tensors QA_q[a][e] = sum_j U[q][j] Gamma~[a][j][e] (+ QB for the mirrored
//                    rows) and the one-body rows (h U D^T)[x], (h^T U D)[x] of the shard: what K1's
//                    epilogue contracts the finished tiles with
//   k_tail_reduce    fixed-order sum of K1's per-slab records into dE/dU rows, partial energy,
//                    one-shot all-reduce over peer memory, optimiser transition (oo_optimize)
//   k_qcontract      T3[x][j][e]  = sum_q U[q][j] * Y[x,q][e]      (oo_transform only)
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
#include "oo_k3.cuh"

namespace oo {

// Is slab (t,q) the one of its pair that gets streamed?
__host__ __device__ inline bool pair_selected(int t, int q) {
  if (t == q) return true;
  return (((t + q) & 1) == 0) ? (t < q) : (t > q);
}

// Closed forms for the pair-symmetric slab list of one row t (slabs ordered by increasing q):
//   q < t selected iff q has parity (t+1)&1;  q == t;  q > t selected iff q has parity t&1.
__host__ __device__ inline int pair_count_below(int t) { return (t + 1) >> 1; }   // selected q < t
__host__ __device__ inline int pair_row_count(int t, int M) {
  return pair_count_below(t) + 1 + ((M - 1 - t) >> 1);
}
// i-th selected q of row t
__host__ __device__ inline int pair_ith_q(int t, int i) {
  const int nb = pair_count_below(t);
  if (i < nb) return ((t + 1) & 1) + 2 * i;
  return t + 2 * (i - nb);
}
// position of the selected slab (t,q) inside row t's list
__host__ __device__ inline int pair_rank(int t, int q) {
  if (q < t) return q >> 1;              // (q + 1 - p) >> 1 with p = parity of q
  return pair_count_below(t) + ((q - t) >> 1);
}

constexpr int QC_ECHUNK = 64;    // e-values per CTA in k_qcontract (each lane owns 2)
constexpr int QC_LANES = QC_ECHUNK / 2;
constexpr int QC_GROUPS = 8;     // the term list is split 8 ways inside the CTA
constexpr int QC_THREADS = QC_LANES * QC_GROUPS;   // 256
constexpr int QC_UNROLL = 8;     // tile loads in flight per thread
constexpr int QC_TBIT = 1 << 30; // term flag: read the transposed tile copy

struct QCParams {
  const double* Y;       // [nslab][Np*Np]
  const double* YT;      // [nslab][Np*Np] transposed tiles (pair-symmetric mode)
  const double* Upad;    // [M][Np] zero-padded U (written by K1's CTA 0)
  double* T3;            // [nrows][Np][Np*Np]
  const int* rowstart;   // pair-symmetric mode: [mloc] first slab index of each row; NULL = dense
  int dense_mirror;      // dense mode only: row x takes the tiles (p,x) for all p of the shard
                         // (partner p, not transposed) instead of its own tiles (x,q)
  const int* done_flag;
  int M, t0, mloc;
  int row0, nrows;       // rows x produced: dense [t0, t0+mloc); pair-symmetric [0, M)
};

static inline size_t qc_smem_bytes(int NT, int M, int mloc) {
  const int Np = NT * 8;
  return (size_t)QC_GROUPS * Np * QC_ECHUNK * sizeof(double) + (size_t)2 * (M + mloc) * sizeof(int);
}

// grid (nrows, ceil(Np^2/64)), block 256.
// Row x:  T3[x][j][e] = sum over the tiles that involve row x of  U[partner][j] * tile[e]
//   own slabs (x,q)   : partner q, tile Y[slab]      (x in this GPU's shard)
//   mirrored (t,x)    : partner t, tile YT[slab]     (pair-symmetric mode, t in the shard, t != x)
// HBM/L2-bound (every tile is read once per orientation): 16-byte loads, 8 in flight per thread.
template <int NT>
__global__ void __launch_bounds__(QC_THREADS, NT <= 2 ? 2 : 1) k_qcontract(const QCParams p) {
  constexpr int Np = NT * 8, Np2 = Np * Np;
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  extern __shared__ double qc_smem[];
  double* red = qc_smem;                                         // [GROUPS][Np][ECHUNK]
  int* s_tile = reinterpret_cast<int*>(red + QC_GROUPS * Np * QC_ECHUNK);
  int* s_urow = s_tile + (p.M + p.mloc);
  const int tid = threadIdx.x, M = p.M;
  const int x = p.row0 + blockIdx.x;
  const bool mine = (x >= p.t0) && (x < p.t0 + p.mloc);

  // ---- term list in closed form (no search, no compaction), all threads ----
  //   own terms      i in [0, n_own)      : slab rowstart[x-t0] + i, partner q_i
  //   mirrored terms t in shard, t != x, slab (t,x) selected: t < x with t = x (mod 2),
  //                                                        t > x with t = x+1 (mod 2)
  const int t1 = p.t0 + p.mloc;
  int n_own = 0, n_lo = 0, n_hi = 0, lo_first = 0, hi_first = 0;
  if (p.rowstart) {
    if (mine) n_own = pair_row_count(x, M);
    const int lo_end = min(x, t1);                       // t in [t0, lo_end), parity x&1
    lo_first = p.t0 + (((x & 1) - (p.t0 & 1)) & 1);
    n_lo = lo_end > lo_first ? (lo_end - lo_first + 1) >> 1 : 0;
    const int hi_beg = max(x + 1, p.t0);                 // t in [hi_beg, t1), parity (x+1)&1
    hi_first = hi_beg + ((((x + 1) & 1) - (hi_beg & 1)) & 1);
    n_hi = t1 > hi_first ? (t1 - hi_first + 1) >> 1 : 0;
  } else if (p.dense_mirror) {
    n_own = p.mloc;                                      // tiles (t0+i, x)
  } else if (mine) {
    n_own = M;
  }
  const int cnt = n_own + n_lo + n_hi;
  for (int i = tid; i < cnt; i += QC_THREADS) {
    int tile, urow;
    if (i < n_own) {
      if (p.rowstart) {
        tile = __ldg(p.rowstart + (x - p.t0)) + i;
        urow = pair_ith_q(x, i);
      } else if (p.dense_mirror) {
        tile = i * M + x;
        urow = p.t0 + i;
      } else {
        tile = (x - p.t0) * M + i;
        urow = i;
      }
    } else {
      const int k = i - n_own;
      const int t = k < n_lo ? lo_first + 2 * k : hi_first + 2 * (k - n_lo);
      tile = (__ldg(p.rowstart + (t - p.t0)) + pair_rank(t, x)) | QC_TBIT;
      urow = t;
    }
    s_tile[i] = tile;
    s_urow[i] = urow;
  }
  __syncthreads();

  const int el = (tid & (QC_LANES - 1)) * 2, grp = tid / QC_LANES;
  const int e = blockIdx.y * QC_ECHUNK + el;       // this thread owns e and e+1 (Np2 is even)
  const bool valid = e < Np2;
  const int ee = valid ? e : 0;
  double acc[Np][2];
#pragma unroll
  for (int j = 0; j < Np; ++j) acc[j][0] = acc[j][1] = 0.0;
  const int per = (cnt + QC_GROUPS - 1) / QC_GROUPS;
  const int n0 = grp * per, n1 = min(cnt, n0 + per);

  auto tile_ptr = [&](int tile) {
    const double* src = (tile & QC_TBIT) ? p.YT : p.Y;
    return reinterpret_cast<const double2*>(src + (size_t)(tile & (QC_TBIT - 1)) * Np2 + ee);
  };
  auto accumulate = [&](const double2 y, int urow) {
    const double2* u2 = reinterpret_cast<const double2*>(p.Upad + (size_t)urow * Np);
#pragma unroll
    for (int j = 0; j < Np / 2; ++j) {
      const double2 u = __ldg(u2 + j);
      acc[2 * j][0] = fma(u.x, y.x, acc[2 * j][0]);
      acc[2 * j][1] = fma(u.x, y.y, acc[2 * j][1]);
      acc[2 * j + 1][0] = fma(u.y, y.x, acc[2 * j + 1][0]);
      acc[2 * j + 1][1] = fma(u.y, y.y, acc[2 * j + 1][1]);
    }
  };
  int n = n0;
#pragma unroll 1
  for (; n + QC_UNROLL <= n1; n += QC_UNROLL) {
    double2 y[QC_UNROLL];
    int ur[QC_UNROLL];
#pragma unroll
    for (int u = 0; u < QC_UNROLL; ++u) {
      y[u] = __ldg(tile_ptr(s_tile[n + u]));
      ur[u] = s_urow[n + u];
    }
#pragma unroll
    for (int u = 0; u < QC_UNROLL; ++u) accumulate(y[u], ur[u]);
  }
  for (; n < n1; ++n) accumulate(__ldg(tile_ptr(s_tile[n])), s_urow[n]);

#pragma unroll
  for (int j = 0; j < Np; ++j)
    *reinterpret_cast<double2*>(red + (grp * Np + j) * QC_ECHUNK + el) =
        make_double2(acc[j][0], acc[j][1]);
  __syncthreads();
  // fixed-order sum over the groups; thread (grp, lane) finishes planes j = grp, grp+8, ...
  if (valid) {
    for (int j = grp; j < Np; j += QC_GROUPS) {
      double2 s = make_double2(0.0, 0.0);
#pragma unroll
      for (int w = 0; w < QC_GROUPS; ++w) {
        const double2 v = *reinterpret_cast<const double2*>(red + (w * Np + j) * QC_ECHUNK + el);
        s.x += v.x;
        s.y += v.y;
      }
      *reinterpret_cast<double2*>(p.T3 + ((size_t)blockIdx.x * Np + j) * Np2 + e) = s;
    }
  }
}

// One-shot all-reduce over NVLink peer memory, fused into the last CTA of k_tail_reduce.
// Every GPU owns a buffer  flags[2][world] | slots[2][world][stride]  that its peers map through
// CUDA IPC.  Evaluation number `seq` (same on all ranks) uses parity seq&1:
//   push   my (gradient | energy) vector into slot [parity][my_rank] of EVERY rank (remote stores;
//          the gradient rows by the CTAs that produce them, the energy by the last CTA)
//   signal flags[parity][my_rank] = seq on every rank (after a system-scope fence)
//   wait   until my own flags[parity][r] >= seq for all r
//   sum    my slots[parity][0..world) in rank order -> identical bits on every rank.
// A rank can be at most one evaluation ahead of the slowest one (it needs everybody's flag to
// finish), so two parities suffice.
constexpr int PEER_MAX = 8;
struct PeerComm {
  double* slots[PEER_MAX];               // rank r's slot area (peer-mapped)
  unsigned long long* flags[PEER_MAX];   // rank r's flag area (peer-mapped)
  int* error_flag;                       // device flag, set on wait time-out (never hang the GPU);
                                         // sticky: every later result is poisoned with NaN
  int* error_flag_host;                  // same flag in mapped host memory (read by the C API)
  unsigned long long timeout_ns;         // how long a rank waits for its peers
  unsigned long long* seq_ptr;           // device-resident evaluation counter (same on all ranks;
                                         // kept on the device so the launch can live in a CUDA graph)
  int rank, world, enabled;
  int stride;                            // doubles per slot (>= M*N+1)
};

constexpr int TAIL_THREADS = 256;

// One finished element of this GPU's vector straight into the slot of every rank (called by the
// CTA that produced it; made visible by that CTA's system-scope fence before it reports in).
__device__ __forceinline__ void peer_push_value(const PeerComm& cm, int idx, double val) {
  const unsigned long long seq = *cm.seq_ptr + 1ull;
  const size_t my_slot = ((size_t)(seq & 1ull) * cm.world + cm.rank) * cm.stride;
  for (int r = 0; r < cm.world; ++r) cm.slots[r][my_slot + idx] = val;
}

// The one-shot all-reduce, executed by all TAIL_THREADS threads of ONE CTA (the last CTA of the
// tail kernel, after it has written the complete local vector buf[0 .. len)).
// `first` = index of the first element this CTA still has to push: 0 pushes the whole vector;
// len - 1 only the energy, when the CTAs that produced the gradient rows have pushed them
// themselves (peer_push_value) -- the 7 x 32 KB of remote stores are then spread over all CTAs of
// the tail kernel instead of being serialised in its last one.
__device__ inline void peer_allreduce_cta(const PeerComm& cm, double* buf, int len, int first = 0) {
  const int tid = threadIdx.x;
  __syncthreads();
  const unsigned long long seq = *cm.seq_ptr + 1ull;
  const int par = (int)(seq & 1ull);
  const size_t my_slot = ((size_t)par * cm.world + cm.rank) * cm.stride;
  for (int idx = first + tid; idx < len; idx += TAIL_THREADS) {
    const double val = __ldcg(buf + idx);
    for (int r = 0; r < cm.world; ++r) cm.slots[r][my_slot + idx] = val;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < cm.world)
    *((volatile unsigned long long*)(cm.flags[tid] + par * cm.world + cm.rank)) = seq;
  if (tid < cm.world) {
    volatile unsigned long long* f = cm.flags[cm.rank] + par * cm.world + tid;
    const unsigned long long t_start = global_timer_ns();
    while (*f < seq) {
      if (global_timer_ns() - t_start > cm.timeout_ns) {   // give up instead of hanging
        *cm.error_flag = 1;
        *((volatile int*)cm.error_flag_host) = 1;
        break;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  // a time-out (now or earlier) means stale slots and ranks that no longer agree bit for bit:
  // poison the result instead of returning a wrong one
  const bool poisoned = *((volatile int*)cm.error_flag) != 0;
  const double* mine_slots = cm.slots[cm.rank] + (size_t)par * cm.world * cm.stride;
  for (int idx = tid; idx < len; idx += TAIL_THREADS) {
    double sum = 0.0;
    for (int r = 0; r < cm.world; ++r) sum += __ldcg(mine_slots + (size_t)r * cm.stride + idx);
    buf[idx] = poisoned ? __longlong_as_double(0x7ff8000000000000ll) : sum;
  }
  __syncthreads();                 // everybody has read *seq_ptr
  if (tid == 0) *cm.seq_ptr = seq;
}

// ---------------------------------------------------------------------------------------------
// Tiles path of the evaluation (N in 25..32, where the fused epilogue of K1 cannot keep up: 2048
// FP64 instructions per slab for two warps): K1 stores the tiles, k_qcontract builds T3 and
// k_tail_row contracts it with the 2-RDM.
// ---------------------------------------------------------------------------------------------
struct TailParams {
  PeerComm comm;
  const double* T3;    // [nrows][Np^3]
  const double* Gp;    // [Np^3][Np]: 2-RDM, a fastest (see k_prepare_gamma)
  const double* U;     // [M][N]
  const double* B1;    // [mloc][N]  one-body rows of the shard
  const double* B12;   // [mloc][N]
  double* out;         // [M*N + 1]: gradient rows + energy (rows outside [row0,row0+nrows) untouched)
  double* rowE;        // [Np/AC][nrows] per-row energy partials
  unsigned int* counter;
  const int* done_flag;
  int M, N, t0, mloc;  // shard: the one-body terms are added for rows in [t0, t0+mloc) only
  int row0, nrows;
  double two_body_grad_factor;  // 4 for the one-pass (V4-symmetric) gradient
  int accumulate;               // 0 (kept for the kernel's generality: out[x] += A[x])
  int do_step;                  // run the optimiser transition in the last CTA
  StepParams step;
};

// Register blocking of the 2-RDM contraction: R rows x AC a-values per CTA (grid.y = Np / AC).
// The 2-RDM is re-read from L2 once per CTA row-group and the T3 rows once per a-chunk, so the
// L2 traffic is nrows/R * |Gp| + Np/AC * |T3|; R*AC accumulators live in registers.
__host__ __device__ constexpr int tail_rows(int NT) { return NT <= 2 ? 4 : 8; }
__host__ __device__ constexpr int tail_ac(int NT) { return NT >= 1 ? 8 : 8; }

// grid (ceil(nrows / R), Np / AC), block 256.  For the R rows x and the AC values a of the CTA:
//   A[x][a] = sum_{j,e} T3[x][j][e] * Gp[a][j][e]                                 (2-RDM contraction)
//   out[x][a] = 4 A[x][a] + B12[x][a] (own rows),   rowE[y][x] = sum_a U[x][a] (A + B1)[x][a]
// and the last CTA adds rowE in fixed order into out[M*N].
template <int NT>
__global__ void __launch_bounds__(TAIL_THREADS) k_tail_row(const TailParams p) {
  constexpr int Np = NT * 8, L = Np * Np * Np, NW = TAIL_THREADS / 32, R = tail_rows(NT),
                AC = tail_ac(NT);
  pdl_launch_dependents();
  pdl_wait();
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  __shared__ double s_part[NW][R][AC];
  __shared__ double s_e[R][AC];
  __shared__ StepSmem sm;
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int xl0 = blockIdx.x * R, a0 = blockIdx.y * AC, N = p.N;

  double acc[R][AC];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int a = 0; a < AC; ++a) acc[r][a] = 0.0;
  const double* tp[R];
#pragma unroll
  for (int r = 0; r < R; ++r) tp[r] = p.T3 + (size_t)min(xl0 + r, p.nrows - 1) * L;
  // each thread takes two consecutive (j,e) positions per step: 16-byte loads everywhere
#pragma unroll 2
  for (int idx = tid * 2; idx < L; idx += TAIL_THREADS * 2) {
    double2 tv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) tv[r] = *reinterpret_cast<const double2*>(tp[r] + idx);
    const double2* g0 = reinterpret_cast<const double2*>(p.Gp + (size_t)idx * Np + a0);
    const double2* g1 = reinterpret_cast<const double2*>(p.Gp + (size_t)(idx + 1) * Np + a0);
#pragma unroll
    for (int a2 = 0; a2 < AC / 2; ++a2) {
      const double2 u = __ldg(g0 + a2), v = __ldg(g1 + a2);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r][2 * a2] = fma(tv[r].x, u.x, fma(tv[r].y, v.x, acc[r][2 * a2]));
        acc[r][2 * a2 + 1] = fma(tv[r].x, u.y, fma(tv[r].y, v.y, acc[r][2 * a2 + 1]));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int a = 0; a < AC; ++a) {
      double v = acc[r][a];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_part[warp][r][a] = v;
    }
  __syncthreads();
  if (tid < R * AC) {
    const int r = tid / AC, al = tid - r * AC, a = a0 + al;
    const int xl = xl0 + r, x = p.row0 + xl;
    double ev = 0.0;
    if (xl < p.nrows && a < N) {
      double av = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) av += s_part[w][r][al];
      const bool mine = (x >= p.t0) && (x < p.t0 + p.mloc);
      double b1 = 0.0, b12 = 0.0;
      if (mine) {
        b1 = p.B1[(size_t)(x - p.t0) * N + a];
        b12 = p.B12[(size_t)(x - p.t0) * N + a];
      }
      if (p.accumulate) p.out[(size_t)x * N + a] += p.two_body_grad_factor * av;
      else p.out[(size_t)x * N + a] = p.two_body_grad_factor * av + b12;
      ev = __ldg(p.U + (size_t)x * N + a) * (av + b1);
    }
    s_e[r][al] = ev;
  }
  __syncthreads();
  if (tid == 0) {
    for (int r = 0; r < R; ++r) {
      if (xl0 + r < p.nrows) {
        double e = 0.0;
        for (int a = 0; a < AC; ++a) e += s_e[r][a];
        p.rowE[(size_t)blockIdx.y * p.nrows + xl0 + r] = e;
      }
    }
    __threadfence();
    const unsigned int prev = atomicAdd(p.counter, 1u);
    is_last = (prev == (unsigned int)(gridDim.x * gridDim.y - 1));
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double v = 0.0;
    for (int i = tid; i < p.nrows * (int)gridDim.y; i += TAIL_THREADS)
      v += ((volatile double*)p.rowE)[i];
    v = block_sum(v, sm.scratch);                  // fixed tree: deterministic
    if (tid == 0) {
      if (!p.accumulate) p.out[(size_t)p.M * N] = v;
      *p.counter = 0u;
    }
    if (p.comm.enabled) peer_allreduce_cta(p.comm, p.out, p.M * N + 1);
    if (p.do_step) {
      __threadfence();
      __syncthreads();
      opt_step_cta<TAIL_THREADS>(p.step, sm);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused evaluation (see oo_k1.cuh): what runs before and after K1.
// ---------------------------------------------------------------------------------------------

// k_prepare_q — everything of an evaluation that depends on U but not on g:
//   z = 0:  QA[p][a][e] = sum_c U[p][c] * G2A[c][a][e]     p in [0, M)
//   z = 1:  QB[pl][a][e] = sum_c U[t0+pl][c] * G2B[c][a][e]  pl in [0, mloc)      (optional)
//   z = 2:  one-body gradient rows of the shard (as k_onebody), one CTA per row
// grid (ceil(Np^3 / 256), row chunks, 3), block 256.  A thread owns one (a, e) position, keeps its
// N coefficients of G2 in registers and walks the rows of its chunk (coalesced 8-byte stores;
// the row of U is a warp-wide broadcast load).
struct PrepParams {
  const double* U;      // [M][N]
  const double* G2A;    // [Np][Np][Np^2]
  const double* G2B;    // [Np][Np][Np^2] or NULL
  double* QA;           // [M][Np][Np^2]
  double* QB;           // [mloc][Np][Np^2]
  const double* h;      // one-body part
  const double* D;
  double* B1;
  double* B12;
  const int* done_flag;
  int M, N, t0, mloc;
  int rows_per_chunk;
  int onebody_only;     // tiles path: only the z = 2 part runs
};

constexpr int PREP_MAX_ROWS = 16;   // rows of U per CTA of k_prepare_q

// Does position idx = a*Np^2 + e1*Np + e0 of a Q matrix carry data (a, e0, e1 < N)?
template <int NT>
__device__ __forceinline__ bool prep_position_live(int idx, int N) {
  constexpr int Np = NT * 8;
  const int a = idx / (Np * Np), e = idx - a * Np * Np, e1 = e / Np, e0 = e - e1 * Np;
  return a < N && e0 < N && e1 < N;
}

template <int NT>
__global__ void __launch_bounds__(256) k_prepare_q(const PrepParams p) {
  constexpr int Np = NT * 8, Np2 = Np * Np, Np3 = Np2 * Np;
  pdl_launch_dependents();
  pdl_wait();                       // U (and the stop flag) come from the previous kernel
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  const int tid = threadIdx.x, N = p.N, M = p.M;
  if (blockIdx.z == 2) {
    // ---- one-body rows: B1 = (h U D^T)[x], B12 = B1 + (h^T U D)[x] ----
    // Pure latency (a 1 x M by M x N product per row): row x and column x of h go to shared
    // memory in one round trip, then every thread walks its slice of q with 8 loads of U in flight.
    const int ob = blockIdx.y * gridDim.x + blockIdx.x;
    if (ob >= p.mloc) return;
    extern __shared__ double s_h[];                    // [2][M]: h[x][:], h[:][x]
    __shared__ double s_r1[256], s_r2[256], s_hu[32], s_htu[32];
    const int x = p.t0 + ob;
    for (int q = tid; q < M; q += 256) {
      s_h[q] = __ldg(p.h + (size_t)x * M + q);
      s_h[M + q] = __ldg(p.h + (size_t)q * M + x);
    }
    __syncthreads();
    const int j = tid % N, part = tid / N, nparts = 256 / N;
    double r1 = 0.0, r2 = 0.0;
    if (part < nparts) {
      int q = part;
      for (; q + 7 * nparts < M; q += 8 * nparts) {
        double u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) u[k] = __ldg(p.U + (size_t)(q + k * nparts) * N + j);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          r1 = fma(s_h[q + k * nparts], u[k], r1);
          r2 = fma(s_h[M + q + k * nparts], u[k], r2);
        }
      }
      for (; q < M; q += nparts) {
        const double u = __ldg(p.U + (size_t)q * N + j);
        r1 = fma(s_h[q], u, r1);
        r2 = fma(s_h[M + q], u, r2);
      }
    }
    s_r1[tid] = r1;
    s_r2[tid] = r2;
    __syncthreads();
    if (tid < N) {
      double hu = 0.0, htu = 0.0;
      for (int w = 0; w < nparts; ++w) {
        hu += s_r1[w * N + tid];
        htu += s_r2[w * N + tid];
      }
      s_hu[tid] = hu;
      s_htu[tid] = htu;
    }
    __syncthreads();
    if (tid < N) {
      double b1 = 0.0, b2 = 0.0;
      for (int jj = 0; jj < N; ++jj) {
        b1 = fma(s_hu[jj], __ldg(p.D + tid * N + jj), b1);    // (hU) D^T
        b2 = fma(s_htu[jj], __ldg(p.D + jj * N + tid), b2);   // (h^T U) D
      }
      p.B1[(size_t)ob * N + tid] = b1;
      p.B12[(size_t)ob * N + tid] = b1 + b2;
    }
    return;
  }
  const bool second = blockIdx.z == 1;
  if (p.onebody_only || (second && p.G2B == nullptr)) return;
  const int nrows = second ? p.mloc : M;
  const int r0 = blockIdx.y * p.rows_per_chunk, r1 = min(nrows, r0 + p.rows_per_chunk);
  if (r0 >= r1) return;
  // the chunk's rows of U, zero-padded to Np columns: read back as 16-byte broadcasts (global
  // broadcast loads made the kernel LSU-bound: N load instructions per row and thread)
  __shared__ double2 s_u[PREP_MAX_ROWS][Np / 2];
  const double* Urow = p.U + (size_t)((second ? p.t0 : 0) + r0) * N;
  for (int i = tid; i < (r1 - r0) * Np; i += 256) {
    const int r = i / Np, c = i - r * Np;
    reinterpret_cast<double*>(&s_u[r][0])[c] = c < N ? __ldg(Urow + (size_t)r * N + c) : 0.0;
  }
  const int idx = blockIdx.x * 256 + tid;             // (a, e) position
  // positions in the padding (a, k or l >= N) are zero from the allocation and stay zero
  const bool valid = idx < Np3 && prep_position_live<NT>(idx, N);
  const double* G2 = second ? p.G2B : p.G2A;
  double coef[Np];
#pragma unroll
  for (int c = 0; c < Np; ++c) coef[c] = valid ? __ldg(G2 + (size_t)c * Np3 + idx) : 0.0;
  __syncthreads();
  if (!valid) return;
  double* out = (second ? p.QB : p.QA) + (size_t)r0 * Np3 + idx;
  for (int r = 0; r < r1 - r0; ++r) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int c2 = 0; c2 < Np / 2; ++c2) {
      const double2 u = s_u[r][c2];
      s0 = fma(u.x, coef[2 * c2], s0);
      s1 = fma(u.y, coef[2 * c2 + 1], s1);
    }
    out[(size_t)r * Np3] = s0 + s1;
  }
}

// k_tail_reduce — what is left of an evaluation after K1: for every row x of U
//   A[x][a] = sum over the streamed slabs that involve row x of their Aslab record, in fixed order
//      own slabs (x, q)      record 0   (x in this GPU's shard)
//      mirror slabs (t, x)   record 1   (t in the shard; pair mode: the selected pairs, t != x;
//                                        generic mode: all t)
//   out[x][a] = factor * A[x][a] + B12[x][a] (own rows),  rowE[x] = sum_a U[x][a] (A' + B1)[x][a]
// then the last CTA adds rowE in fixed order into out[M*N], runs the one-shot all-reduce over
// peer memory when attached and, inside oo_optimize, the optimiser transition itself (k_step's
// body), so that one iteration is three launches: k_prepare_q, k1_half_transform, k_tail_reduce.
// grid M rows (pair / generic) or mloc rows (dense), block 256 = (256 / Np) term groups x Np.
struct TailReduceParams {
  PeerComm comm;
  const double* Aslab;   // [nslab][2][Np]
  const int* rowstart;   // pair mode: [mloc] first slab of each row; NULL = dense slab order
  const double* U;
  const double* B1;
  const double* B12;
  double* out;
  double* rowE;
  unsigned int* counter;
  const int* done_flag;
  int M, N, t0, mloc;
  int row0, nrows;
  int mirror_mode;       // 0 none (dense V4), 1 pair-selected transposed tiles, 2 all t (generic)
  int energy_mirror;     // 1: the mirror records enter the energy as well (V4: A is complete)
  double grad_factor;    // 4 for the one-pass V4 gradient, 1 per generic slot pair
  int accumulate;        // generic second pass: out += A, energy untouched
  int do_step;           // 1: run the optimiser transition in the last CTA; 2: its small-problem
                         // variant (M*N <= STEP_SMALL_MN)
  int step_smem;         // the launch carries M*N doubles of dynamic shared memory for V
  StepParams step;
};

template <int NT>
__global__ void __launch_bounds__(TAIL_THREADS) k_tail_reduce(const TailReduceParams p) {
  constexpr int Np = NT * 8, G = TAIL_THREADS / Np;
  pdl_launch_dependents();
  pdl_wait();                       // Aslab comes from K1
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  __shared__ double s_part[G][Np];
  __shared__ double s_e[Np];
  __shared__ bool is_last;
  __shared__ StepSmem sm;
  __shared__ StepSmallSmem ss;
  const int tid = threadIdx.x, a = tid % Np, grp = tid / Np, N = p.N, M = p.M;
  // small problems: every CTA loads what the optimiser transition will need (any of them may be
  // the last one), in the shadow of the row reduction; OO_NO_STEP_FUSION / large M*N: off
  const bool small_step = p.do_step == 2;
  if (small_step) opt_step_small_prefetch<TAIL_THREADS>(p.step, ss);
  // fused all-reduce: every CTA pushes its own row to the peers when the launch covers all rows
  // (dense first-index shards leave the other rows to the bulk push of the last CTA: zeros)
  const bool row_push = p.comm.enabled && p.nrows == M;
  const bool act = grp < G;                 // Np = 24: the last 16 threads have no group
  const int x = p.row0 + blockIdx.x;
  const bool mine = (x >= p.t0) && (x < p.t0 + p.mloc);
  const int t1 = p.t0 + p.mloc;

  // term lists in closed form (same enumeration as k_qcontract)
  int n_own = 0, n_lo = 0, n_hi = 0, lo_first = 0, hi_first = 0, own_base = 0;
  if (mine) {
    n_own = p.rowstart ? pair_row_count(x, M) : M;
    own_base = p.rowstart ? __ldg(p.rowstart + (x - p.t0)) : (x - p.t0) * M;
  }
  if (p.mirror_mode == 1) {
    const int lo_end = min(x, t1);                       // t in [t0, lo_end), parity x&1
    lo_first = p.t0 + (((x & 1) - (p.t0 & 1)) & 1);
    n_lo = lo_end > lo_first ? (lo_end - lo_first + 1) >> 1 : 0;
    const int hi_beg = max(x + 1, p.t0);                 // t in [hi_beg, t1), parity (x+1)&1
    hi_first = hi_beg + ((((x + 1) & 1) - (hi_beg & 1)) & 1);
    n_hi = t1 > hi_first ? (t1 - hi_first + 1) >> 1 : 0;
  } else if (p.mirror_mode == 2) {
    n_lo = p.mloc;                                       // slabs (t0 + i, x)
  }
  const int n_mir = n_lo + n_hi;
  double s_own = 0.0, s_mir = 0.0;
  for (int i = grp; act && i < n_own; i += G)
    s_own += __ldcg(p.Aslab + ((size_t)(own_base + i) * 2) * Np + a);
  for (int i = grp; act && i < n_mir; i += G) {
    int slab;
    if (p.mirror_mode == 1) {
      const int t = i < n_lo ? lo_first + 2 * i : hi_first + 2 * (i - n_lo);
      slab = __ldg(p.rowstart + (t - p.t0)) + pair_rank(t, x);
    } else {
      slab = i * M + x;
    }
    s_mir += __ldcg(p.Aslab + ((size_t)slab * 2 + 1) * Np + a);
  }
  // fixed-order sum over the groups: own part first, then the mirrored part
  if (act) s_part[grp][a] = s_own;
  __syncthreads();
  double a_own = 0.0, a_mir = 0.0;
  if (tid < Np) {
#pragma unroll
    for (int w = 0; w < G; ++w) a_own += s_part[w][tid];
  }
  __syncthreads();
  if (act) s_part[grp][a] = s_mir;
  __syncthreads();
  if (tid < Np) {
#pragma unroll
    for (int w = 0; w < G; ++w) a_mir += s_part[w][tid];
    double ev = 0.0;
    if (tid < N) {
      const double av = a_own + a_mir;
      double b1 = 0.0, b12 = 0.0;
      if (mine) {
        b1 = p.B1[(size_t)(x - p.t0) * N + tid];
        b12 = p.B12[(size_t)(x - p.t0) * N + tid];
      }
      double gval = p.grad_factor * av;
      if (p.accumulate) gval += p.out[(size_t)x * N + tid];
      else gval += b12;
      p.out[(size_t)x * N + tid] = gval;
      if (row_push) peer_push_value(p.comm, x * N + tid, gval);
      ev = __ldg(p.U + (size_t)x * N + tid) * ((p.energy_mirror ? av : a_own) + b1);
    }
    s_e[tid] = ev;
  }
  __syncthreads();
  if (tid == 0) {
    double e = 0.0;
    for (int i = 0; i < Np; ++i) e += s_e[i];
    p.rowE[blockIdx.x] = e;
    if (row_push) __threadfence_system();   // the remote stores of this row, before reporting in
    else __threadfence();
    const unsigned int prev = atomicAdd(p.counter, 1u);
    is_last = (prev == (unsigned int)(gridDim.x - 1));
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  {
    double v = 0.0;
    for (int i = tid; i < p.nrows; i += TAIL_THREADS) v += ((volatile double*)p.rowE)[i];
    v = block_sum(v, sm.scratch);                  // fixed tree: deterministic
    if (tid == 0) {
      if (!p.accumulate) p.out[(size_t)M * N] = v;
      *p.counter = 0u;
    }
  }
  if (p.comm.enabled) peer_allreduce_cta(p.comm, p.out, M * N + 1, row_push ? M * N : 0);
  if (p.do_step) {
    __threadfence();
    __syncthreads();
    extern __shared__ double s_step_v[];      // M*N doubles when the launch reserved them
    if (small_step) opt_step_small_cta<TAIL_THREADS>(p.step, sm, ss);
    else opt_step_cta<TAIL_THREADS>(p.step, sm, p.step_smem ? s_step_v : nullptr);
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
  __shared__ double scratch[33];
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
