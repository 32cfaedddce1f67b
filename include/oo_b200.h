/*
 * oo_b200.h — C ABI of liboo_b200.so: the B200 (sm_100a) implementation of the orbital-optimisation
 * inner loop of OptOrbVQE (energy E(U) and gradient dE/dU of the partial-unitary projection
 * optimiser, Stiefel retraction, Barzilai-Borwein step).
 *
 * The reference (JoelHBierman/electronic-structure-orbital-optimization) is pure Python and has no
 * FFI; each entry point below names the reference code it replaces, paths relative to
 * electronic_structure_algorithms/orbital_optimization/ :
 *   pupo.py = partial_unitary_projection_optimizer.py,  base.py = base_opt_orb_solver.py,
 *   eig.py  = opt_orb_eigensolver.py.
 *
 * Conventions
 *   - All matrices are FP64, row-major.  "dev" pointers are CUDA device pointers on the context's
 *     device, "host" pointers are ordinary host memory.  No torch types cross this boundary.
 *   - Spatial-orbital picture: M orbitals in the large basis, N in the active space (N <= 32).
 *       h  [M][M]           one-body integrals (one spin block of the reference's tensor)
 *       g  [Mloc][M][M][M]  two-body integrals, "physicist" index order of base.py:90, rows
 *                           t0 .. t0+Mloc-1 of the first index (the GPU's shard; Mloc = M when
 *                           unsharded)
 *       D  [N][N]           spin-summed, state-weighted 1-RDM
 *       G  [N][N][N][N]     spin-summed, state-weighted 2-RDM
 *       U  [M][N]           partial unitary
 *     E(U) = sum h_pq U_pi U_qj D_ij + sum g_pqrs U_pi U_qj U_rk U_sl G_ijkl   (base.py:554-563)
 *   - Every function returns 0 on success, a negative oo_status otherwise; oo_last_error() gives a
 *     thread-local description.  Nothing aborts the process.  There is no CPU fallback: a missing
 *     or non-sm_100 device is an error.
 */
#ifndef OO_B200_H
#define OO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oo_ctx oo_ctx;

enum oo_status {
  OO_OK = 0,
  OO_ERR_INVALID = -1,   /* bad argument / shape                                  */
  OO_ERR_CUDA = -2,      /* CUDA runtime or driver error                          */
  OO_ERR_STATE = -3,     /* call sequence error (integrals or RDMs not set ...)   */
  OO_ERR_NCCL = -4,      /* NCCL not loadable or a collective failed              */
  OO_ERR_UNSUPPORTED = -5,
  OO_ERR_NUMERIC = -6    /* non-finite value met during the optimisation          */
};

/* oo_set_integrals flags */
#define OO_G_V4_SYMMETRIC 1u /* caller asserts g[pqrs]=g[qpsr]=g[rspq]=g[srqp] (holds for real
                                orbitals); enables the one-pass analytic gradient              */

#define OO_G_PAIR_PACKED 2u  /* with OO_G_V4_SYMMETRIC: g_dev holds only the slabs that the
                                pair-symmetric mode streams (see oo_pair_slab_list), half the
                                memory of the dense shard                                      */

/* ---- library / device ---------------------------------------------------------------------- */
const char* oo_last_error(void);
const char* oo_version(void);
/* Number of visible CUDA devices with compute capability 10.x; <0 on error. */
int oo_device_count(void);

/* ---- context ------------------------------------------------------------------------------- */
/* One context per GPU (one process per GPU).  The context owns all workspaces; input tensors
 * stay owned by the caller and must outlive their use.  Shard = rows [t0, t0+mloc) of g's first
 * index; pass t0=0, mloc=M for a single GPU.
 * Replaces: the implicit torch device state behind PartialUnitaryProjectionOptimizer(device=...)
 * (pupo.py:15-48). */
int oo_create(int device, int M, int N, int t0, int mloc, oo_ctx** out);
int oo_destroy(oo_ctx* ctx);
/* Use an existing cudaStream_t (e.g. torch's current stream) for all work; NULL = own stream. */
int oo_set_stream(oo_ctx* ctx, void* cuda_stream);
int oo_synchronize(oo_ctx* ctx);

/* ---- inputs -------------------------------------------------------------------------------- */
/* Registers (does not copy) h_dev [M][M] and g_dev [mloc][M][M][M].  M must be even (TMA global
 * strides are multiples of 16 bytes).  Replaces the .to(device) shuttling of the integral tensors,
 * opt_orb_minimum_eigensolver.py:219-222. */
int oo_set_integrals(oo_ctx* ctx, const double* h_dev, const double* g_dev, unsigned flags);
/* Pair-packed storage (flag OO_G_PAIR_PACKED).  Because g[t,q,r,s] = g[q,t,s,r], one slab of every
 * pair {(t,q),(q,t)} carries all the information; the packed shard keeps exactly the slabs the
 * pair-symmetric mode streams, as [slab][r][s] in streaming order (t ascending, then q ascending).
 * 8*M*M*count bytes instead of 8*M*M*M*mloc: M=400 takes 102.4 GB (fits one 180 GB GPU) instead
 * of 204.8 GB.  Kernel traffic and arithmetic are identical to dense storage.
 *   oo_pair_slab_list: host-only; writes (t,q) of the i-th stored slab to tq_host[2i], [2i+1]
 *     (tq_host may be NULL) and returns the slab count for rows [t0, t0+mloc), <0 on error.
 *   oo_pack_pair_slabs: gathers the packed shard from a dense shard g_dense_dev [mloc][M][M][M]
 *     on the context's stream (for callers that start from the reference's dense tensor,
 *     base.py:89-90). */
int oo_pair_slab_list(int M, int t0, int mloc, int* tq_host, int capacity);
int oo_pack_pair_slabs(oo_ctx* ctx, const double* g_dense_dev, double* g_packed_dev);
/* Two-body tensors WITHOUT V4 symmetry: registers this GPU's shard of g, g_dev [mloc][M][M][M]
 * (rows t0.. of the FIRST index), and the same rows of the pair-transposed tensor,
 * g_pt_dev[r][s][p][q] = g[p][q][r][s] for r in [t0, t0+mloc) (i.e. the shard of g's THIRD index,
 * stored with that index first).  Every evaluation then makes two dense passes (one per tensor)
 * and builds dE/dU from the four index-slot terms, with the 2-RDM used as given (no
 * symmetrisation): rows t and q of dE/dU come from the first pass, rows r and s from the second,
 * every GPU contributes partial rows for all of U and the usual all-reduce completes them.  Four
 * times the work of the symmetric path; completes the contract of compute_rotated_energy for
 * arbitrary real tensors (base.py:534-563, pupo.py:85-103). */
int oo_set_integrals_generic(oo_ctx* ctx, const double* h_dev, const double* g_dev,
                             const double* g_pt_dev);
/* Max |g - g∘pi| over the three V4 permutations and max |g| of a FULL (unsharded) device tensor
 * g_dev[M][M][M][M]; out_host[0]=max asymmetry, out_host[1]=max |g|. */
int oo_check_v4_symmetry(int device, const double* g_dev, int M, double* out_host);
/* Spatial RDMs D_dev [N][N], G_dev [N][N][N][N]; symmetrised/padded copies are made.
 * Replaces the per-state RDM arguments of base.py:534-538 / the state loop of eig.py:149-169
 * (weights are folded in by the caller: E is linear in the RDMs). */
int oo_set_rdms(oo_ctx* ctx, const double* D_dev, const double* G_dev);

/* Spin-orbital ingest of the reference's two-body tensor g_spin_dev [P]^4, P = 2M (alpha block
 * first; base.py:89-90).  Because W = block_diag(U,U) (base.py:549) the energy only couples
 * matching spin blocks; block id b = 8*s0 + 4*s1 + 2*s2 + s3.  Finds the non-zero blocks
 * (|.| > rtol * max|g|), verifies that they are identical (restricted integrals) and copies the
 * common spatial tensor to g_sp_out_dev [M]^4.  *block_mask gets one bit per non-zero block;
 * stats_host[0] = max|g|, stats_host[1] = max deviation between non-zero blocks.
 * OO_ERR_UNSUPPORTED when the blocks differ (unrestricted integrals). */
int oo_ingest_spin_g(int device, const double* g_spin_dev, int M, double rtol,
                     double* g_sp_out_dev, unsigned* block_mask, double* stats_host);
/* The same for one shard: only rows [t0, t0+mloc) of the first index of the spatial block are
 * extracted (and compared across the non-zero spin blocks), into g_out_dev [mloc][Mpad][Mpad][Mpad]
 * with Mpad >= M (Mpad = M+1 pads an odd M to the even extent oo_create needs; the padding must
 * be zero on entry).  A rank of a multi-GPU run never materialises the M^4 spatial tensor. */
int oo_ingest_spin_g_rows(int device, const double* g_spin_dev, int M, double rtol, int t0,
                          int mloc, int Mpad, double* g_out_dev, unsigned* block_mask,
                          double* stats_host);
/* RDMs straight from the reference's spin-orbital tensors (base.py:362-532 layout): nstates (<=8)
 * device tensors D_n [2N][2N], G_n [2N]^4 given as HOST arrays of device pointers, host weights
 * (NULL = 1): D~ = sum_n w_n (D_n[aa] + D_n[bb]), G~ = sum_n w_n sum_{b in block_mask} G_n[block b]
 * (eig.py:149-169: the weighted energy sum is linear in the RDMs), then as oo_set_rdms. */
int oo_set_rdms_spin(oo_ctx* ctx, const double* const* D_spin_dev, const double* const* G_spin_dev,
                     const double* weights_host, int nstates, unsigned block_mask);

/* Pair-symmetric slab mode (default: enabled).  A V4-symmetric tensor satisfies
 * g[t,q,r,s] = g[q,t,s,r], so the half-transformed tiles obey Y[q,t] = Y[t,q]^T: only one slab of
 * every pair {(t,q),(q,t)} is streamed from HBM (checkerboard choice: ~M/2 slabs per row, shards
 * stay balanced) and used for both rows.  With several GPUs every rank then contributes partial
 * gradient rows for ALL of U, completed by the same all-reduce.  enable=0 streams every slab of the
 * shard (dense mode: out_dev rows outside the shard are zero). */
int oo_set_pair_symmetry(oo_ctx* ctx, int enable);
/* Number of M x M slabs one evaluation streams from HBM on this GPU (algorithmic bytes =
 * 8*M*M*slabs). */
int oo_streamed_slabs(oo_ctx* ctx);

/* ---- evaluation ---------------------------------------------------------------------------- */
/* Enqueue one evaluation at U_dev [M][N].  out_dev [M*N+1] receives this GPU's partial dE/dU
 * (pair-symmetric mode: partial values for every row; dense mode: the shard's rows, other rows
 * are zeroed) followed by the GPU's partial energy;
 * the sum over GPUs is E(U), dE/dU.  With one GPU no reduction is needed.  Asynchronous on the
 * context stream.
 * Replaces base.py:534-582 (compute_rotated_energy) + pupo.py:85-103 (autograd gradient). */
int oo_energy_grad(oo_ctx* ctx, const double* U_dev, double* out_dev);
/* As oo_energy_grad, followed by the sum over all GPUs: through the one-shot all-reduce fused into
 * the last kernel (NVLink peer memory, oo_peer_attach) when available, else through NCCL
 * (oo_comm_init).  out_dev then holds E(U) and dE/dU on every rank, bit-identical across ranks in
 * the fused mode (fixed rank-order summation). */
int oo_energy_grad_allreduce(oo_ctx* ctx, const double* U_dev, double* out_dev);
/* Same through host buffers: H2D of U, evaluation, all-reduce when a communicator is attached,
 * D2H of E and dE/dU, synchronous.  This is the reference-facing call used for end-to-end timing. */
int oo_energy_grad_host(oo_ctx* ctx, const double* U_host, double* E_host, double* grad_host);
/* The same call split in two so that host-buffer evaluations can be pipelined (two slots, 0 and
 * 1): oo_eval_submit copies U_host into pinned memory and enqueues H2D (own copy stream),
 * evaluation (+ all-reduce) and D2H (own copy stream) without waiting; oo_eval_wait blocks until
 * the slot's result has landed and hands it out (grad_host may be NULL).  Submitting slot s+1 before
 * waiting for slot s overlaps the copies and the host work of one evaluation with the kernels of
 * the next; the finite-difference gradient (pupo.py:105-127: 2*M*N energies) and bench.py's
 * end-to-end leg use it.  oo_energy_grad_host = submit + wait on slot 0. */
int oo_eval_submit(oo_ctx* ctx, const double* U_host, int slot);
int oo_eval_wait(oo_ctx* ctx, int slot, double* E_host, double* grad_host);
/* Rotated integrals h' [N][N], g' [N][N][N][N] (this shard's partial sum over the first index).
 * Replaces the tensor part of get_rotated_hamiltonian, base.py:597-604, and
 * opt_orb_mcvqe.py:90-98. */
int oo_transform(oo_ctx* ctx, const double* U_dev, double* h_rot_dev, double* g_rot_dev);

/* ---- retraction / BB step ------------------------------------------------------------------ */
/* U_out = V (V^T V)^(-1/2).  Replaces orth(), pupo.py:70-83 and base.py:614-626. */
int oo_orth(oo_ctx* ctx, const double* V_dev, double* U_out_dev);
/* compute_updated_partial_unitary, pupo.py:129-159.  alpha_io_dev: BB step size in/out. */
int oo_bb_update(oo_ctx* ctx, int iteration, const double* U_cur_dev, const double* U_prev_dev,
                 const double* G_cur_dev, const double* G_prev_dev, double* alpha_io_dev,
                 double* U_new_dev);

/* ---- the whole inner loop ------------------------------------------------------------------ */
/* compute_optimal_rotation, pupo.py:161-350, entirely on the device: U_io_host [M][N] is the
 * initial partial unitary on entry and the final one on exit; *E_final = the reference's return
 * value P4_array[0]; E_hist_host[k] = f(U_k) for k < hist_cap (callback replay);
 * *n_iter = final iteration_number. */
/* Optional live callback of oo_optimize, the reference's callback(iteration, energy)
 * (pupo.py:29-30): invoked on the calling host thread, in order and with the reference's
 * arguments, at most one chunk (4 iterations) after the device produced the value, while the
 * device keeps iterating.  NULL disables it. */
typedef void (*oo_callback_t)(int iteration, double energy, void* user);
int oo_set_callback(oo_ctx* ctx, oo_callback_t cb, void* user);
/* May be called from inside a callback: the running oo_optimize stops as if the reference's loop
 * had ended at the iteration the device has reached (it returns OO_OK with that iterate).  Used by
 * the Python mirror to propagate an exception raised by the user's callback. */
int oo_request_stop(oo_ctx* ctx);
int oo_optimize(oo_ctx* ctx, double* U_io_host, double bb0, double tol, int maxiter, double decay,
                double* E_hist_host, int hist_cap, int* n_iter, double* E_final,
                double* bb_final);

/* ---- multi-GPU (one process per GPU) -------------------------------------------------------- */
/* 128-byte NCCL unique id, to be created on rank 0 and broadcast by the host framework. */
int oo_nccl_unique_id(void* id128_host);
int oo_comm_init(oo_ctx* ctx, const void* id128_host, int rank, int world);
/* In-place NCCL sum all-reduce of count doubles on the context stream. */
int oo_allreduce(oo_ctx* ctx, double* buf_dev, size_t count);
/* Fused all-reduce over NVLink peer memory (one process per GPU, CUDA IPC).  oo_peer_export
 * allocates this GPU's exchange buffer and returns its 64-byte cudaIpcMemHandle_t; the host
 * framework all-gathers the handles; oo_peer_attach(handles[world][64]) maps the peers' buffers.
 * From then on oo_energy_grad_allreduce / oo_energy_grad_host / oo_optimize do the all-reduce
 * inside the evaluation's last kernel: the last CTA pushes the (M*N+1)-double vector into every
 * peer, signals a flag, waits for the peers' flags and sums the contributions in rank order.
 * oo_peer_status: 1 = attached and healthy, 0 = not attached, <0 = a wait timed out. */
int oo_peer_export(oo_ctx* ctx, void* handle64_host);
int oo_peer_attach(oo_ctx* ctx, const void* handles_host, int rank, int world);
int oo_peer_status(oo_ctx* ctx);
/* How long a rank waits for its peers inside the fused all-reduce (default 20 s, or the
 * environment variable OO_PEER_TIMEOUT_MS).  On a time-out the result of that and of every later
 * evaluation is NaN on the device and oo_optimize / oo_energy_grad_host / oo_eval_wait /
 * oo_synchronize return OO_ERR_NCCL: a stalled peer can never turn into a silently wrong
 * gradient.  Host callbacks run on the enqueueing thread and can stall the peers for as long
 * as they take. */
int oo_set_peer_timeout_ms(oo_ctx* ctx, double milliseconds);

/* ---- measurement helpers -------------------------------------------------------------------- */
/* Device time (ms) of the kernels of the last oo_energy_grad call, measured with CUDA events on
 * the context stream: [0] K1 half-transform with the fused 2-RDM contraction, [1] k_prepare_q
 * (Q tensors + one-body rows), [2] k_tail_reduce (row sums, energy, all-reduce), [3] reserved
 * (0), [4] whole evaluation.  Requires oo_set_timing(ctx,1) beforehand. */
int oo_set_timing(oo_ctx* ctx, int enable);
int oo_last_timing(oo_ctx* ctx, float* ms5_host);
/* Number of kernels launched by this context since creation. */
long long oo_launch_count(oo_ctx* ctx);
/* Telemetry of the last oo_optimize: Newton-Schulz iterations summed over all retractions and the
 * number of retractions that fell back to the Jacobi eigensolver. */
int oo_retraction_stats(oo_ctx* ctx, int* newton_schulz_iterations, int* jacobi_fallbacks);
/* Measured FP64 peaks of the device: out_host[0] = DMMA.8x8x4 TFLOP/s (register-resident loop),
 * out_host[1] = DFMA TFLOP/s, out_host[2] = streaming read GB/s (LDG.128 sum over `bytes`). */
int oo_measure_peaks(int device, size_t bytes, double* out_host);

#ifdef __cplusplus
}
#endif
#endif /* OO_B200_H */
