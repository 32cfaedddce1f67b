// K1 — fused half-transform of the two-electron integrals (the M^4*N leading term).
//
// Replaces the first two pairwise contractions of the reference's 6-operand einsum
//   torch.einsum('pqrs,pi,qj,rk,sl,ijkl', g, W, W, W, W, Gamma)
// (electronic_structure_algorithms/orbital_optimization/base_opt_orb_solver.py:558-563) in the
// spatial-orbital picture:  for every slab (t,q) of the ERI tensor g[t][q][r][s]
//
//      Y_tq[k][l] = sum_{r,s} U[r][k] * g[t][q][r][s] * U[s][l]         (N x N per slab)
//
// Each slab (M x M doubles, contiguous) is streamed from HBM exactly once by TMA into a multi-stage
// shared-memory ring (128B-swizzled boxes of 256 rows x 16 columns), and contracted twice on the
// FP64 tensor cores (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) without the intermediate ever leaving
// registers:
//      Z^T[l][r]  = sum_s U[s][l] * g[r][s]      A = U^T (smem), B = g tile (smem, swizzled)
//      Y^T[l][k] += sum_r Z^T[l][r] * U[r][k]    A = Z^T accumulator fragments (register reuse:
//                                                 the 8x8 C fragment of m8n8k4 is two 8x4 A
//                                                 fragments with a permuted k order), B = U^T (smem)
// Work decomposition: persistent CTAs (one per SM), slabs dealt round-robin; inside a CTA one TMA
// producer warp, three epilogue warps (per-slab reduction + fused 2-RDM contraction, off the MMA
// critical path) and 8 consumer warps, every consumer warp owns up to 4 row-blocks (8 rows each) of
// the current 256-row pass and all of its columns, so the second contraction costs a fixed N/M
// fraction of the first.
//
// Fused 2-RDM contraction (energy / gradient evaluations).  The rows of A = dE2/dU / 4 are
//      A[t][a] = sum_q <Y_tq, QA_q[a]>   +   sum_{t'} <Y_t't, QB_t'[a]>   (second sum: pair mode)
// with QA_q[a][e] = sum_j U[q][j] Gamma~[a][j][e] (k_prepare_q; depends on U and the 2-RDM only).
// The epilogue warps therefore never store the N x N tiles: warp C sums the consumers' partial
// tiles, warp A takes the dot products of the finished tile with QA_q (row t of A), warp B with
// QB_t (row q of A: the transposed tile of the pair partner, the transposition folded into QB's
// layout), and they write 2 x Np doubles per slab (Aslab).  The tiles, the q-contraction and the T3 tensor disappear from the evaluation;
// what is left after K1 is a fixed-order sum of Aslab records per row (k_tail_reduce).
// Tile mode (p.QA == NULL; oo_transform only): warp C stores Y and its transpose.
#pragma once
#include "oo_common.cuh"

namespace oo {

constexpr int K1_NWARP = 8;                        // consumer warps per CTA
constexpr int K1_RB = 4;                           // 8-row blocks per consumer warp per pass
constexpr int K1_ROWS = K1_NWARP * K1_RB * 8;      // slab rows per pass = TMA box height (256)
constexpr int K1_KC = 16;                          // slab columns per stage (128 B swizzle span)
constexpr int K1_STAGE_BYTES = K1_ROWS * K1_KC * 8;  // 32 KiB
constexpr int K1_THREADS = (K1_NWARP + 4) * 32;    // + 1 TMA producer warp + 3 epilogue warps
constexpr int K1_BAR_FULL = 1;                     // named barrier: per-warp partial tiles written
constexpr int K1_BAR_FREE = 2;                     // named barrier: partial-tile buffer reusable
constexpr int K1_BAR_FOLD = 3;                     // first of the two-warp hand-over barriers (ids 3..12)
constexpr int K1_BAR_UT = 13;                      // U^T built (all warps but the producer)
constexpr int K1_BAR_COUNT = (K1_NWARP + 1) * 32;  // consumers + the summing epilogue warp

struct K1Params {
  const double* U;       // [M][N] row-major partial unitary
  double* Y;             // [nslab][Np][Np], element (l,k) of slab = Y_tq[k][l]
  double* YT;            // optional [nslab][Np][Np] transposed tiles (pair-symmetric mode) or NULL
  double* Upad;          // optional [M][Np] zero-padded copy of U written by CTA 0
  const int* done_flag;  // optional: non-zero => kernel is a no-op (optimiser already stopped)
  int M, N;
  const int* slab_coord; // optional: slab i lives at tensor coordinate slab_coord[i] = tl*M + q
                         // (pair-symmetric mode: only one of (t,q)/(q,t) is streamed); NULL = i
  int nslab;             // number of (t,q) slabs this GPU streams (dense: Mloc * M)
  int nstage;            // TMA ring depth
  int stage_tx_bytes;    // bytes one TMA box delivers (box rows x 16 columns; <= K1_STAGE_BYTES)
  int Mk;                // M rounded up to a multiple of K1_KC
  int upitch;            // row pitch (doubles) of the transposed U copy in smem: Mk + 8
  int l2_hints;          // bit 0: stream g with the L2 evict-first policy (it is read exactly once);
                         // bit 1: store the Y / YT tiles evict-last so that the q-contraction finds
                         // them in L2 (used when Y + YT are a small part of L2)
  int npart;             // partial-tile buffers at slab end: 8 (one per warp), 4 or 2 (warps are
                         // folded into them in fixed order; frees smem for one more TMA stage)
  // ---- fused 2-RDM contraction (QA != NULL) ----
  const double* QA;      // [M][Np][Np*Np]    QA_q[a][e], rows of all partner orbitals q
  const double* QB;      // [mloc][Np][Np*Np] QB_t[a][e], rows of this GPU's shard (NULL: no second dot)
  double* Aslab;         // [nslab][2][Np]: <tile, QA_q[a]> and <tile, QB_t[a]> of every streamed slab
  const int* slab_tq;    // slab i is (tl, q) with slab_tq[i] = tl*M + q; NULL = i (dense order)
};

// Dynamic shared memory needed by k1_half_transform<NT> (before 1024 B alignment slack).
static inline size_t k1_smem_bytes(int NT, int Mk, int nstage, int npart = K1_NWARP) {
  const int Np = NT * 8;
  size_t b = (size_t)nstage * K1_STAGE_BYTES;           // TMA ring
  b += (size_t)Np * (Mk + 8) * sizeof(double);          // Ut
  b += (size_t)npart * Np * Np * sizeof(double);        // partial-tile buffers (8, 4 or 2)
  b += (size_t)2 * Np * Np * sizeof(double);            // finished tiles, double buffered
  b += (size_t)(2 * nstage + 4) * sizeof(uint64_t);     // full/empty + tile full/free mbarriers
  return b + 1024;                                      // alignment slack
}

template <int NT>
struct K1Frag {
  double g[K1_RB][2];  // B operand: g[row][s], two consecutive k4 steps
  double u[NT][2];     // A operand: U[s][l]
};

// 2 k4-steps x NACT row-blocks x NT column tiles of DMMA on one fragment set.  NACT = number of
// this warp's row-blocks that lie inside the slab in the current pass (always a prefix).  It is
// a template parameter reached through a warp-uniform switch: a predicated-off DMMA still
// occupies the tensor pipe (measured: profiles/r01_ncu_summary.md), a branch does not.
template <int NT, int NACT>
__device__ __forceinline__ void k1_mma_chunk_n(double (&acc)[K1_RB][NT][2], const K1Frag<NT>& f) {
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int rb = 0; rb < NACT; ++rb)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        dmma884(acc[rb][nt][0], acc[rb][nt][1], f.u[nt][j], f.g[rb][j]);
}

// End of a pass: Y^T[l][k] += sum_r Z^T[l][r] * U[r][k] over the warp's NACT row-blocks, then
// clear the Z^T accumulators.
template <int NT, int NACT>
__device__ __forceinline__ void k1_second_gemm_n(double (&yacc)[NT][NT][2],
                                                 double (&acc)[K1_RB][NT][2], uint32_t ub0,
                                                 uint32_t u_nt_stride) {
#pragma unroll
  for (int rb = 0; rb < NACT; ++rb) {
    const uint32_t ub = ub0 + (uint32_t)(rb * K1_NWARP * 8 * 8);  // 8 rows per block, NWARP apart
    // all B fragments of the row-block first, then the DMMAs ordered so that two updates of the
    // same accumulator are NT*NT issues apart (back-to-back dependent DMMAs stall on the result)
    double b[NT][2];
#pragma unroll
    for (int nk = 0; nk < NT; ++nk) {
      b[nk][0] = lds64(ub + nk * u_nt_stride);       // U[row0 + c    ][nk*8+g]
      b[nk][1] = lds64(ub + nk * u_nt_stride + 32);  // U[row0 + c + 4][nk*8+g]
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int nk = 0; nk < NT; ++nk)
#pragma unroll
        for (int nl = 0; nl < NT; ++nl)
          dmma884(yacc[nl][nk][0], yacc[nl][nk][1], acc[rb][nl][j], b[nk][j]);
  }
#pragma unroll
  for (int rb = 0; rb < K1_RB; ++rb)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[rb][nt][0] = acc[rb][nt][1] = 0.0;
}

// Ring position and per-lane shared-memory addresses of one consumer warp.
struct K1Cons {
  uint32_t g_addr[2];   // this lane's 16 B of row-block `warp` in stage 0, k8 halves 0 / 1
  uint32_t u_addr;      // this lane's A fragment of U^T at column 0
  uint32_t u_nt_stride; // bytes between column tiles of U^T
  uint32_t full_base, empty_base;
  int stage, nstage;
  uint32_t phase;
};

template <int NT>
__device__ __forceinline__ void k1_load_frag(K1Frag<NT>& f, const K1Cons& s, int stage,
                                             uint32_t ucol, int half) {
  const uint32_t sb = s.g_addr[half] + (uint32_t)stage * K1_STAGE_BYTES;
#pragma unroll
  for (int rb = 0; rb < K1_RB; ++rb)
    lds128(f.g[rb][0], f.g[rb][1], sb + (uint32_t)(rb * K1_NWARP * 1024));
  const uint32_t ub = s.u_addr + ucol + (uint32_t)(half * 64);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) lds128(f.u[nt][0], f.u[nt][1], ub + nt * s.u_nt_stride);
}

// One 256-row pass over a slab: nkc chunks of 16 columns through the TMA ring, then GEMM 2.
// NACT (this warp's row-blocks inside the slab in this pass) is a template parameter selected
// once per pass: a predicated-off DMMA still occupies the tensor pipe (measured:
// profiles/r01_ncu_summary.md) and a per-chunk switch costs an indirect branch per 24 DMMAs.
// f0 holds the first k8 half of the next chunk on entry and on exit (software pipeline across
// pass and slab boundaries); `last` = nothing follows this pass.
template <int NT, int NACT>
__device__ __forceinline__ void k1_pass(double (&acc)[K1_RB][NT][2], double (&yacc)[NT][NT][2],
                                        K1Frag<NT>& f0, K1Frag<NT>& f1, K1Cons& s, int nkc,
                                        bool last, uint32_t ub2, int lane) {
  uint32_t ucol = 0;   // byte offset of the chunk's first column inside a row of U^T
  for (int kc = 0; kc < nkc; ++kc) {
    k1_load_frag<NT>(f1, s, s.stage, ucol, 1);
    k1_mma_chunk_n<NT, NACT>(acc, f0);

    int ns = s.stage + 1;
    uint32_t nph = s.phase;
    if (ns == s.nstage) {
      ns = 0;
      nph ^= 1u;
    }
    const bool wrap = kc + 1 == nkc;
    const uint32_t nucol = wrap ? 0u : ucol + (uint32_t)(K1_KC * 8);
    if (!(wrap && last)) {
      mbar_wait(s.full_base + 8u * ns, nph);
      k1_load_frag<NT>(f0, s, ns, nucol, 0);
    }

    k1_mma_chunk_n<NT, NACT>(acc, f1);

    // All shared-memory reads of the stage are complete (their values fed the MMAs above).
    __syncwarp();
    if (lane == 0) mbar_arrive(s.empty_base + 8u * s.stage);
    s.stage = ns;
    s.phase = nph;
    ucol = nucol;
  }
  // Y^T[l][k] += sum_r Z^T[l][r] * U[r][k] over this warp's rows of the pass
  k1_second_gemm_n<NT, NACT>(yacc, acc, ub2, s.u_nt_stride);
}

// <tile, Q[a]> for a = 0 .. Np-1: the epilogue's dot products of one finished N x N tile (held as
// NCH 16-byte chunks per lane) with one Q matrix [Np][Np*Np] in L2; the Np results go to out[].
// FP64 instructions are the scarce resource here: the two consumer warps that share the
// sub-partition keep its FP64 pipe busy with DMMAs, and every DFMA / DADD of this warp waits for a
// gap (ncu: stall_math on ~all of them, ~85 cycles each).  Hence
//   * one DFMA pair per (a, chunk) and nothing else in the loop,
//   * a transposing shuffle reduction: 8 accumulators are reduced with 4+2+1+1+1 = 9 DADDs
//     instead of 8 x 5 (each exchange step halves the number of live values per lane),
//   * loads in groups of 8, two groups in flight.
// Fixed shuffle tree => deterministic.
// q0 = the first load group (a = 0..7, chunk 0), issued by the caller BEFORE it waits for the tile:
// the slab's (t, q) is known in advance, so one L2 round trip leaves the per-slab critical chain.
template <int NT>
__device__ __forceinline__ void k1_epilogue_dots(const double2 (&tile)[NT * NT], const double* Q,
                                                 int lane, uint64_t keep, double* out,
                                                 const double2 (&q0)[8]) {
  constexpr int Np = NT * 8, Np2 = Np * Np, NCH = Np2 / 64, NAG = Np / 8;
  constexpr int DEPTH = 2;
  static_assert(NCH == NT * NT, "chunks per lane");
  // group (ag, i): a = 8*ag .. 8*ag+7, chunk i.  NT <= 3: one flat, fully unrolled sequence of
  // NAG*NCH groups; NT = 4 (64 groups): the a-group loop stays a run-time loop (code size)
  constexpr int OUTER = NT <= 3 ? 1 : NAG;            // run-time iterations
  constexpr int INNER = NT <= 3 ? NAG * NCH : NCH;    // unrolled groups per iteration
  const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0, up4 = (lane & 4) != 0;
#pragma unroll 1
  for (int o = 0; o < OUTER; ++o) {
    const double* Qo = Q + (size_t)o * 8 * Np2;       // NT = 4: this iteration's 8 rows of Q
    double2 qv[DEPTH][8];
    double acc[8];
    static_assert(DEPTH == 2, "the caller prefetches exactly one group");
    if (o == 0) {
#pragma unroll
      for (int a = 0; a < 8; ++a) qv[0][a] = q0[a];
    } else {
#pragma unroll
      for (int a = 0; a < 8; ++a) qv[0][a] = ldg128_hint(Qo + (size_t)a * Np2, keep);
    }
#pragma unroll
    for (int g = 0; g < INNER; ++g) {
      if (g + DEPTH - 1 < INNER) {
        const int gn = g + DEPTH - 1;
        const int ag = NT <= 3 ? gn / NCH : 0, i = gn % NCH;
#pragma unroll
        for (int a = 0; a < 8; ++a)
          qv[gn % DEPTH][a] = ldg128_hint(Qo + (size_t)(ag * 8 + a) * Np2 + 64 * i, keep);
      }
      const int ag = NT <= 3 ? g / NCH : o, i = g % NCH;
      if (i == 0) {
#pragma unroll
        for (int a = 0; a < 8; ++a)
          acc[a] = fma(tile[0].x, qv[g % DEPTH][a].x, tile[0].y * qv[g % DEPTH][a].y);
      } else {
#pragma unroll
        for (int a = 0; a < 8; ++a)
          acc[a] = fma(tile[i].x, qv[g % DEPTH][a].x, fma(tile[i].y, qv[g % DEPTH][a].y, acc[a]));
      }
      if (i == NCH - 1) {
        // transposing reduction over the 32 lanes: after the steps 16, 8, 4 a lane holds ONE
        // value, the partial sum for a = 4*bit4 + 2*bit3 + bit2 of its lane id; steps 2, 1 finish
        double v4[4], v2[2], v;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double keepv = up16 ? acc[k + 4] : acc[k], send = up16 ? acc[k] : acc[k + 4];
          v4[k] = keepv + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const double keepv = up8 ? v4[k + 2] : v4[k], send = up8 ? v4[k] : v4[k + 2];
          v2[k] = keepv + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
          const double keepv = up4 ? v2[1] : v2[0], send = up4 ? v2[0] : v2[1];
          v = keepv + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if ((lane & 3) == 0) out[ag * 8 + (lane >> 2)] = v;
      }
    }
  }
}

// 12 warps: the register cap is 65536 / 384 = 168 per thread (ptxas -v: NT = 1: 89, NT = 2: 125, NT = 3: 168 without spills,
// NT = 4: 168 with ~0.9 KB of spills -- it still reaches 29 TFLOP/s).
template <int NT>
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_half_transform(const __grid_constant__ CUtensorMap tmap, const K1Params p) {
  constexpr int Np = NT * 8;
  pdl_launch_dependents();

  extern __shared__ uint8_t k1_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(k1_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  double* Ut = reinterpret_cast<double*>(smem + (size_t)p.nstage * K1_STAGE_BYTES);
  double* Ypart = Ut + (size_t)Np * p.upitch;
  double* Tbuf = Ypart + p.npart * Np * Np;              // [2][Np*Np] finished tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(Tbuf + 2 * Np * Np);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t stage_base = smem_u32(smem);
  const uint32_t full_base = smem_u32(bars);
  const uint32_t empty_base = full_base + 8u * p.nstage;
  const uint32_t tfull_base = empty_base + 8u * p.nstage;   // [2] tile buffer filled (1 arrival)
  const uint32_t tfree_base = tfull_base + 16u;             // [2] tile buffer read (2 arrivals)

  if (tid == 0) {
    for (int s = 0; s < p.nstage; ++s) {
      mbar_init(full_base + 8u * s, 1);
      mbar_init(empty_base + 8u * s, K1_NWARP);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_base + 8u * b, 1);
      mbar_init(tfree_base + 8u * b, 2);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmap);
  }
  __syncthreads();   // barrier initialisation visible (before any dependency wait: cheap)

  const int npass = (p.M + K1_ROWS - 1) / K1_ROWS;
  const int nkc = p.Mk / K1_KC;
  const int nrb_total = (p.M + 7) / 8;
  int nmine = 0;
  if ((int)blockIdx.x < p.nslab) nmine = (p.nslab - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;

  if (warp == K1_NWARP) {
    // ------------------------------ TMA producer ------------------------------
    // The ERI tensor and the slab table do not depend on the previous kernel, so the ring is
    // filled BEFORE the dependency wait: the first slabs are in flight while the other warps
    // still wait for U / the Q tensors and build U^T (about 2 us per launch, which is 8 % of an
    // iteration of the small configs).
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      int stage = 0, issued = 0;
      uint32_t phase = 0;
      bool checked = p.done_flag == nullptr;     // plain evaluations: nothing to check
      auto stopped = [&]() {
        // the optimiser has already stopped: nobody will consume the ring, but the CTA must not
        // exit while bulk copies into its shared memory are still in flight
        pdl_wait();
        checked = true;
        if (*p.done_flag == 0) return false;
        for (int s2 = 0; s2 < min(issued, p.nstage); ++s2) mbar_wait(full_base + 8u * s2, 0u);
        return true;
      };
      for (int slab = blockIdx.x; slab < p.nslab; slab += gridDim.x) {
        const int coord = p.slab_coord ? __ldg(p.slab_coord + slab) : slab;
        for (int pass = 0; pass < npass; ++pass) {
          for (int kc = 0; kc < nkc; ++kc) {
            if (!checked && issued == p.nstage && stopped()) return;
            mbar_wait(empty_base + 8u * stage, phase ^ 1u);
            mbar_arrive_expect_tx(full_base + 8u * stage, (uint32_t)p.stage_tx_bytes);
            if (p.l2_hints & 1)
              tma_load_3d_hint(stage_base + (uint32_t)stage * K1_STAGE_BYTES, &tmap, kc * K1_KC,
                               pass * K1_ROWS, coord, full_base + 8u * stage, pol);
            else
              tma_load_3d(stage_base + (uint32_t)stage * K1_STAGE_BYTES, &tmap, kc * K1_KC,
                          pass * K1_ROWS, coord, full_base + 8u * stage);
            ++issued;
            if (++stage == p.nstage) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
      if (!checked) stopped();
    }
    return;
  }

  // all other warps: U, the Q tensors and the stop flag are outputs of the previous kernels
  pdl_wait();
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  {
    // Transposed, zero-padded copy of U: Ut[l][s] = U[s][l]  (read row-major, i.e. coalesced:
    // consecutive threads take consecutive l of one row s); 11 warps, the producer is busy
    const int ctid = tid < K1_NWARP * 32 ? tid : tid - 32, cnt = K1_THREADS - 32;
    for (int idx = ctid; idx < Np * p.upitch; idx += cnt) {
      const int s = idx / Np, l = idx - s * Np;
      Ut[l * p.upitch + s] = (l < p.N && s < p.M) ? __ldg(p.U + (size_t)s * p.N + l) : 0.0;
    }
    // Side product for the q-contraction (tile mode): the zero-padded row-major copy Upad[M][Np].
    if (blockIdx.x == 0 && p.Upad != nullptr) {
      for (int idx = ctid; idx < p.M * Np; idx += cnt) {
        const int s = idx / Np, l = idx - s * Np;
        p.Upad[idx] = (l < p.N) ? __ldg(p.U + (size_t)s * p.N + l) : 0.0;
      }
    }
    named_bar_sync(K1_BAR_UT, K1_THREADS - 32);
  }

  if (warp == K1_NWARP + 1) {
    // ------------------------------ epilogue warp C: the summer ------------------------------
    // Takes the per-slab reduction off the consumers' critical path: they only drop their
    // partial tiles into shared memory and go on with the next slab.  This warp sums the
    // partials in fixed order (deterministic), releases the partial-tile buffers at once and
    // hands the finished tile to the two dot-product warps through a double-buffered tile
    // (fused mode) or stores it and its transpose (tile mode).
    constexpr int NCH = Np * Np / 64;                      // 16-byte chunks of the tile per lane
    named_bar_arrive(K1_BAR_FREE, K1_BAR_COUNT);           // the buffers start out free
    const uint64_t keep = l2_policy_evict_last();
    int it = 0;
    for (int slab = blockIdx.x; slab < p.nslab; slab += gridDim.x, ++it) {
      named_bar_sync(K1_BAR_FULL, K1_BAR_COUNT);
      double2 tile[NCH];
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        double2 s2 = make_double2(0.0, 0.0);
        for (int w = 0; w < p.npart; ++w) {
          const double2 v =
              *reinterpret_cast<const double2*>(Ypart + w * Np * Np + 64 * i + 2 * lane);
          s2.x += v.x;
          s2.y += v.y;
        }
        tile[i] = s2;
      }
      named_bar_arrive(K1_BAR_FREE, K1_BAR_COUNT);
      if (p.QA != nullptr) {
        const int b = it & 1, n = it >> 1;                 // n-th use of tile buffer b
        if (n > 0) mbar_wait(tfree_base + 8u * b, (uint32_t)((n - 1) & 1));
        double2* dst = reinterpret_cast<double2*>(Tbuf + b * Np * Np) + lane;
#pragma unroll
        for (int i = 0; i < NCH; ++i) dst[32 * i] = tile[i];
        __syncwarp();
        if (lane == 0) mbar_arrive(tfull_base + 8u * b);
      } else {
        // tile mode (oo_transform): the tile and, in pair-symmetric mode, its transpose, so that
        // the q-contraction reads both orientations with unit stride
        double* out = p.Y + (size_t)slab * Np * Np;
        double* outT = p.YT ? p.YT + (size_t)slab * Np * Np : nullptr;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int e = 64 * i + 2 * lane + h;
            const double v = h ? tile[i].y : tile[i].x;
            if (p.l2_hints & 2) {
              st_global_hint(out + e, v, keep);
              if (outT) st_global_hint(outT + (e % Np) * Np + e / Np, v, keep);
            } else {
              out[e] = v;
              if (outT) outT[(e % Np) * Np + e / Np] = v;
            }
          }
        }
      }
    }
    return;
  }

  if (warp > K1_NWARP + 1) {
    // ------------------------- epilogue warps A, B: the dot products -------------------------
    // fused mode only: warp A contracts the finished tile with QA_q (row t of A), warp B with
    // QB_t (row q); 2 x Np doubles per slab leave the SM instead of the tile.
    if (p.QA == nullptr) return;
    constexpr int NCH = Np * Np / 64;
    const bool second = warp == K1_NWARP + 3;
    const bool active = !second || p.QB != nullptr;
    const uint64_t keep = l2_policy_evict_last();
    int it = 0;
    for (int slab = blockIdx.x; slab < p.nslab; slab += gridDim.x, ++it) {
      const int tq = p.slab_tq ? __ldg(p.slab_tq + slab) : slab;
      const int b = it & 1, n = it >> 1;
      const int tl = tq / p.M, q = tq - tl * p.M;
      const double* Q = (second ? p.QB + (size_t)tl * Np * Np * Np : p.QA + (size_t)q * Np * Np * Np) +
                        2 * lane;
      double2 q0[8];
      if (active) {
#pragma unroll
        for (int a = 0; a < 8; ++a) q0[a] = ldg128_hint(Q + (size_t)a * Np * Np, keep);
      }
      mbar_wait(tfull_base + 8u * b, (uint32_t)(n & 1));
      double2 tile[NCH];
      if (active) {
        const double2* src = reinterpret_cast<const double2*>(Tbuf + b * Np * Np) + lane;
#pragma unroll
        for (int i = 0; i < NCH; ++i) tile[i] = src[32 * i];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(tfree_base + 8u * b);
      if (!active) continue;
      k1_epilogue_dots<NT>(tile, Q, lane, keep,
                           p.Aslab + ((size_t)slab * 2 + (second ? 1 : 0)) * Np, q0);
    }
    return;
  }

  // -------------------------------- consumers ---------------------------------
  const int g = lane >> 2, c = lane & 3;
  // MMA column n=g of the B operand maps to slab row (8*block + rho): rows whose swizzle phases
  // differ in bit 2 sit in the same quarter-warp, which makes every LDS.128 conflict-free.
  const int rho = (g >> 1) | ((g & 1) << 2);
  // Byte offset of this lane's 16 B (two doubles: columns 2c, 2c+1 of the k8 half) inside an
  // 8-row block of a swizzled stage; half h in {0,1} selects columns [8h, 8h+8).
  K1Cons cs;
  cs.g_addr[0] = stage_base + (uint32_t)(warp * 1024 + rho * 128 + (((0 * 4 + c) ^ rho) << 4));
  cs.g_addr[1] = stage_base + (uint32_t)(warp * 1024 + rho * 128 + (((1 * 4 + c) ^ rho) << 4));
  const uint32_t ut_base = smem_u32(Ut);
  // A operand for GEMM 1: lane reads Ut[nt*8+g][col + 2c .. 2c+1]
  cs.u_addr = ut_base + (uint32_t)((g * p.upitch + 2 * c) * 8);
  // B operand for GEMM 2: lane reads Ut[nt*8+g][row + c] and [row + c + 4]
  const uint32_t u_addr_b = ut_base + (uint32_t)((g * p.upitch + c + warp * 8) * 8);
  cs.u_nt_stride = (uint32_t)(8 * p.upitch * 8);
  cs.full_base = full_base;
  cs.empty_base = empty_base;
  cs.stage = 0;
  cs.nstage = p.nstage;
  cs.phase = 0;

  double acc[K1_RB][NT][2];   // Z^T fragments of the current pass
  double yacc[NT][NT][2];     // Y^T fragments of the current slab
#pragma unroll
  for (int rb = 0; rb < K1_RB; ++rb)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[rb][nt][0] = acc[rb][nt][1] = 0.0;
#pragma unroll
  for (int a = 0; a < NT; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) yacc[a][b][0] = yacc[a][b][1] = 0.0;

  K1Frag<NT> f0, f1;
  if (nmine > 0) {
    mbar_wait(full_base, 0);
    k1_load_frag<NT>(f0, cs, 0, 0u, 0);
  }

  static_assert(K1_RB == 4, "the switch below enumerates 0..4 active row-blocks");
  for (int slab = blockIdx.x; slab < p.nslab; slab += gridDim.x) {
    const bool last_slab = slab + (int)gridDim.x >= p.nslab;
    for (int pass = 0; pass < npass; ++pass) {
      // this warp's row-blocks rb*NWARP+warp inside the slab form a prefix of length nact (0..RB)
      const int nrb_pass = min(K1_NWARP * K1_RB, nrb_total - pass * K1_NWARP * K1_RB);
      const int nact = max(0, min(K1_RB, (nrb_pass - warp + K1_NWARP - 1) / K1_NWARP));
      const bool last = last_slab && pass == npass - 1;
      const uint32_t ub2 = u_addr_b + (uint32_t)(pass * K1_ROWS * 8);
      switch (nact) {
        case 4: k1_pass<NT, 4>(acc, yacc, f0, f1, cs, nkc, last, ub2, lane); break;
        case 3: k1_pass<NT, 3>(acc, yacc, f0, f1, cs, nkc, last, ub2, lane); break;
        case 2: k1_pass<NT, 2>(acc, yacc, f0, f1, cs, nkc, last, ub2, lane); break;
        case 1: k1_pass<NT, 1>(acc, yacc, f0, f1, cs, nkc, last, ub2, lane); break;
        default: k1_pass<NT, 0>(acc, yacc, f0, f1, cs, nkc, last, ub2, lane); break;
      }
    }
    // ---- end of slab: hand the per-warp partial tile to the epilogue warp ----
    named_bar_sync(K1_BAR_FREE, K1_BAR_COUNT);   // previous slab's tile has been consumed
    // warps w, w+npart, w+2*npart, ... share buffer w % npart and add their tiles to it one
    // after the other (fixed order => deterministic).  Each hand-over is a two-warp
    // arrive/sync pair on its own named barrier, so nobody else waits.
    const int buf = warp % p.npart, turn = warp / p.npart;
    double* mine = Ypart + buf * Np * Np;
    if (turn > 0) named_bar_sync(K1_BAR_FOLD + buf * 3 + (turn - 1), 64);
#pragma unroll
    for (int nl = 0; nl < NT; ++nl)
#pragma unroll
      for (int nk = 0; nk < NT; ++nk) {
        double2* dst = reinterpret_cast<double2*>(mine + (nl * 8 + g) * Np + nk * 8 + 2 * c);
        double2 v = make_double2(yacc[nl][nk][0], yacc[nl][nk][1]);
        if (turn > 0) {
          const double2 o = *dst;
          v.x += o.x;
          v.y += o.y;
        }
        *dst = v;
        yacc[nl][nk][0] = yacc[nl][nk][1] = 0.0;
      }
    if ((turn + 1) * p.npart < K1_NWARP) named_bar_arrive(K1_BAR_FOLD + buf * 3 + turn, 64);
    named_bar_arrive(K1_BAR_FULL, K1_BAR_COUNT);
  }
}

}  // namespace oo
