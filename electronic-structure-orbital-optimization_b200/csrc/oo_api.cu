// liboo_b200.so — C ABI (include/oo_b200.h) over the sm_100a kernels.
// Host side only orchestrates: context, workspaces, TMA descriptor, launches, NCCL (dlopen'ed).
#include "../../include/oo_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "oo_bench.cuh"
#include "oo_common.cuh"
#include "oo_ingest.cuh"
#include "oo_k1.cuh"
#include "oo_k2.cuh"
#include "oo_k3.cuh"

using namespace oo;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU_TRY(expr)                                                                     \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return fail(OO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                  __FILE__, __LINE__);                                                   \
  } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL, resolved at run time so the library loads on machines without it
// ------------------------------------------------------------------------------------------------
namespace {
struct NcclUniqueId128 {
  char internal[128];
};
typedef int (*nccl_get_unique_id_t)(NcclUniqueId128*);
typedef int (*nccl_comm_init_rank_t)(void**, int, NcclUniqueId128, int);
typedef int (*nccl_all_reduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_comm_destroy_t)(void*);
typedef const char* (*nccl_get_error_string_t)(int);
constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64
constexpr int kNcclSum = 0;      // ncclRedOp_t::ncclSum

struct NcclApi {
  void* handle = nullptr;
  nccl_get_unique_id_t get_unique_id = nullptr;
  nccl_comm_init_rank_t comm_init_rank = nullptr;
  nccl_all_reduce_t all_reduce = nullptr;
  nccl_comm_destroy_t comm_destroy = nullptr;
  nccl_get_error_string_t get_error_string = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return OO_OK;
  const char* cand[] = {getenv("OO_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* c : cand) {
    if (!c || !*c) continue;
    h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(OO_ERR_NCCL, "cannot dlopen libnccl.so.2 (set OO_NCCL_LIB): %s", dlerror());
  NcclApi a;
  a.handle = h;
  a.get_unique_id = (nccl_get_unique_id_t)dlsym(h, "ncclGetUniqueId");
  a.comm_init_rank = (nccl_comm_init_rank_t)dlsym(h, "ncclCommInitRank");
  a.all_reduce = (nccl_all_reduce_t)dlsym(h, "ncclAllReduce");
  a.comm_destroy = (nccl_comm_destroy_t)dlsym(h, "ncclCommDestroy");
  a.get_error_string = (nccl_get_error_string_t)dlsym(h, "ncclGetErrorString");
  if (!a.get_unique_id || !a.comm_init_rank || !a.all_reduce || !a.comm_destroy ||
      !a.get_error_string)
    return fail(OO_ERR_NCCL, "libnccl is missing a required symbol");
  g_nccl = a;
  return OO_OK;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct oo_ctx {
  int device = 0, M = 0, N = 0, NT = 0, Np = 0, t0 = 0, mloc = 0;
  int num_sms = 0, nstage = 0, Mk = 0, npart = 8, box_rows = 256;
  size_t k1_smem = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  const double* h = nullptr;
  const double* g = nullptr;
  const double* g2 = nullptr;          // generic mode: pair-transposed tensor g2[r,s,p,q] = g[p,q,r,s]
  bool generic = false;                // no V4 symmetry: two dense passes, four gradient slots
  bool packed = false;                 // g holds only the pair-selected slabs, in streaming order
  alignas(64) CUtensorMap tmap2;
  double* G2[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // k_prepare_gamma2 kinds
  double *QA = nullptr, *QB = nullptr, *Aslab = nullptr;   // fused evaluation (oo_k1.cuh)
  bool step_fusable = true;            // OO_NO_STEP_FUSION=1: k_step stays a separate launch
  bool tiles_eval = false;             // evaluation through stored tiles (N in 25..32; OO_EVAL_PATH)
  double* Gp = nullptr;                // tiles path: 2-RDM, V4-averaged, [Np^3][Np]
  unsigned gflags = 0;
  bool have_ints = false, have_rdms = false;
  alignas(64) CUtensorMap tmap;
  // workspaces (device)
  double *Y = nullptr, *T3 = nullptr, *D = nullptr, *rowE = nullptr, *out = nullptr, *Ucur = nullptr, *Uprev = nullptr,
         *Gprev = nullptr, *E_hist = nullptr,
         *YT = nullptr, *Upad = nullptr, *B1 = nullptr, *B12 = nullptr, *Gtmp = nullptr;
  int hist_cap = 0;
  unsigned int* counter = nullptr;
  // pair-symmetric slab selection (see oo_k2.cuh)
  int* slab_coord = nullptr;    // [nsel] tensor coordinate tl*M+q of the i-th streamed slab
  int* rowstart = nullptr;      // [mloc] index of the first streamed slab of each row
  int nsel = 0;
  bool pair_sym = true;
  OptState* state = nullptr;
  // pinned host staging
  double* pin = nullptr;        // M*N+1 doubles
  OptState* pin_state = nullptr;  // 2 slots
  cudaEvent_t poll_ev[2] = {nullptr, nullptr};
  // NCCL
  void* comm = nullptr;
  int rank = 0, world = 1;
  // fused peer-memory all-reduce (CUDA IPC)
  void* peer_base = nullptr;            // my flags|slots allocation
  void* peer_map[PEER_MAX] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool peer_on = false;
  int peer_stride = 0;
  unsigned long long* peer_seq_dev = nullptr;
  int* peer_err = nullptr;              // device flag (sticky): a peer wait timed out
  volatile int* peer_err_host = nullptr;   // the same flag in mapped host memory
  int* peer_err_host_dev = nullptr;        // its device address
  unsigned long long peer_timeout_ns = 20000000000ull;   // OO_PEER_TIMEOUT_MS / oo_set_peer_timeout_ms
  // pipelined host-buffer evaluations (oo_eval_submit / oo_eval_wait): two slots
  double* slot_pin_u[2] = {nullptr, nullptr};
  double* slot_pin_out[2] = {nullptr, nullptr};
  double* slot_u[2] = {nullptr, nullptr};
  double* slot_out[2] = {nullptr, nullptr};
  cudaEvent_t slot_h2d[2] = {nullptr, nullptr}, slot_eval[2] = {nullptr, nullptr},
              slot_done[2] = {nullptr, nullptr};
  bool slot_busy[2] = {false, false};
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  volatile int stop_requested = 0;      // oo_request_stop (from a callback)
  // CUDA graph of one chunk of optimiser transitions (oo_optimize)
  cudaGraphExec_t chunk_graph = nullptr;
  const void* graph_key[10] = {nullptr, nullptr, nullptr, nullptr, nullptr,
                               nullptr, nullptr, nullptr, nullptr, nullptr};
  long long graph_kernels = 0;
  // timing
  bool timing = false;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float last_ms[5] = {0, 0, 0, 0, 0};
  long long launches = 0;
  int last_ns_iters = 0, last_jacobi_calls = 0;   // telemetry of the last oo_optimize
  int force_jacobi = 0;   // OO_FORCE_JACOBI=1: retraction through the eigensolver path only
  // live callback of oo_optimize (pupo.py:193-194, 226-227, 260-261, 312-313)
  oo_callback_t cb = nullptr;
  void* cb_user = nullptr;
  double* pin_hist = nullptr;   // 2 slots x chunk energies
};

namespace {

typedef CUresult (*encode_tiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(encode_tiled_t* fn) {
  static encode_tiled_t cached = nullptr;
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p)
      return fail(OO_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    cached = (encode_tiled_t)p;
  }
  *fn = cached;
  return OO_OK;
}

// 3-D view of the shard: (s: M, r: M, slab: nslab), box 16 x 256 x 1, 128B swizzle, zero fill.
// nslab = mloc*M for dense storage, the number of stored slabs for pair-packed storage.
int build_tmap(oo_ctx* c, const double* base, CUtensorMap* out_map, size_t nslab = 0) {
  if (nslab == 0) nslab = (size_t)c->mloc * c->M;
  encode_tiled_t enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  const cuuint64_t dims[3] = {(cuuint64_t)c->M, (cuuint64_t)c->M, (cuuint64_t)nslab};
  const cuuint64_t strides[2] = {(cuuint64_t)c->M * 8, (cuuint64_t)c->M * c->M * 8};
  // box height: the 256 rows of a pass, or just the slab's rows (rounded up to a row-block) when
  // the whole slab is smaller -- TMA fills (and the mbarrier counts) the full box, zeros included
  const cuuint32_t box[3] = {K1_KC, (cuuint32_t)c->box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OO_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return OO_OK;
}

// fused = true: the evaluation path (dot products with the Q tensors in the epilogue, Aslab out);
// fused = false: tile mode (Y / YT out), used by oo_transform only.
void fill_peer_comm(oo_ctx* c, PeerComm& cm) {
  cm.enabled = 1;
  cm.rank = c->rank;
  cm.world = c->world;
  cm.stride = c->peer_stride;
  cm.seq_ptr = c->peer_seq_dev;
  cm.error_flag = c->peer_err;
  cm.error_flag_host = c->peer_err_host_dev;
  cm.timeout_ns = c->peer_timeout_ns;
  for (int r = 0; r < c->world; ++r) {
    cm.flags[r] = (unsigned long long*)c->peer_map[r];
    cm.slots[r] = (double*)((char*)c->peer_map[r] + 256);
  }
}

// Kernel launch with the programmatic-dependent-launch attribute (see oo_common.cuh): the three
// kernels of an evaluation and the evaluations of an optimiser chunk overlap their launch and
// prologue with the drain of their predecessor.  Plain launch while per-kernel timing is on (the
// events between the kernels need hard boundaries) or with OO_NO_PDL=1.
template <typename Kernel, typename... Args>
cudaError_t launch_chain(oo_ctx* c, Kernel kernel, dim3 grid, dim3 block, size_t smem,
                         Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool off = getenv("OO_NO_PDL") != nullptr;
  if (!off && !c->timing) {
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int NT>
int launch_k1_t(oo_ctx* c, const double* U, const int* done_flag, bool second_tensor, bool fused) {
  K1Params p;
  memset(&p, 0, sizeof p);
  p.U = U;
  const bool pair = c->pair_sym && !c->generic;
  p.done_flag = done_flag;
  p.M = c->M;
  p.N = c->N;
  p.slab_coord = (pair && !c->packed) ? c->slab_coord : nullptr;   // packed: slab i is stored i-th
  p.nslab = pair ? c->nsel : c->mloc * c->M;
  p.nstage = c->nstage;
  p.stage_tx_bytes = c->box_rows * K1_KC * (int)sizeof(double);
  p.Mk = c->Mk;
  p.upitch = c->Mk + 8;
  p.npart = c->npart;
  if (fused) {
    p.QA = c->QA;
    p.QB = (pair || c->generic) ? c->QB : nullptr;
    p.Aslab = c->Aslab;
    p.slab_tq = pair ? c->slab_coord : nullptr;
    // the ERI stream is read exactly once: evict-first, so that the Q tensors (re-read by every
    // slab) keep their place in L2.  OO_L2_HINTS=<mask> overrides (experiments).
    static const char* env = getenv("OO_L2_HINTS");
    p.l2_hints = 1;
    if (env && *env >= '0' && *env <= '3') p.l2_hints = *env - '0';
  } else {
    p.Y = c->Y;
    p.YT = pair ? c->YT : nullptr;
    p.Upad = c->Upad;
    const size_t tile_bytes = (size_t)p.nslab * c->Np * c->Np * sizeof(double) * (pair ? 2 : 1);
    p.l2_hints = tile_bytes <= ((size_t)48 << 20) ? 3 : 0;
  }
  static bool attr_set[8] = {false, false, false, false, false, false, false, false};
  if (!attr_set[c->device & 7]) {
    CU_TRY(cudaFuncSetAttribute(k1_half_transform<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                227 * 1024));
    attr_set[c->device & 7] = true;
  }
  const int grid = std::min(c->num_sms, p.nslab);
  if (fused) {
    CU_TRY(launch_chain(c, k1_half_transform<NT>, dim3(grid), dim3(K1_THREADS), c->k1_smem,
                        second_tensor ? c->tmap2 : c->tmap, p));
  } else {
    k1_half_transform<NT><<<grid, K1_THREADS, c->k1_smem, c->stream>>>(
        second_tensor ? c->tmap2 : c->tmap, p);
    CU_TRY(cudaGetLastError());
  }
  c->launches++;
  return OO_OK;
}

int launch_k1(oo_ctx* c, const double* U, const int* done_flag, bool second_tensor, bool fused) {
  switch (c->NT) {
    case 1: return launch_k1_t<1>(c, U, done_flag, second_tensor, fused);
    case 2: return launch_k1_t<2>(c, U, done_flag, second_tensor, fused);
    case 3: return launch_k1_t<3>(c, U, done_flag, second_tensor, fused);
    case 4: return launch_k1_t<4>(c, U, done_flag, second_tensor, fused);
  }
  return fail(OO_ERR_INVALID, "unsupported N");
}

// Q tensors (QA for every orbital, QB for the shard's rows) and the one-body rows.
template <int NT>
int launch_prep_t(oo_ctx* c, const double* U, const int* done_flag, int kindA, int kindB,
                  bool onebody_only) {
  constexpr int Np3 = NT * 8 * NT * 8 * NT * 8;
  PrepParams pp;
  pp.U = U;
  pp.G2A = c->G2[kindA];
  pp.G2B = kindB >= 0 ? c->G2[kindB] : nullptr;
  pp.QA = c->QA;
  pp.QB = c->QB;
  pp.h = c->h;
  pp.D = c->D;
  pp.B1 = c->B1;
  pp.B12 = c->B12;
  pp.done_flag = done_flag;
  pp.M = c->M;
  pp.N = c->N;
  pp.t0 = c->t0;
  pp.mloc = c->mloc;
  pp.onebody_only = onebody_only ? 1 : 0;
  const int nbx = (Np3 + 255) / 256;
  // rows of U per CTA: enough CTAs to fill the GPU twice, at most PREP_MAX_ROWS (shared-memory
  // copy of the rows), as many as possible otherwise (every CTA re-reads its coefficients)
  int chunks = std::max(1, std::min(c->M, (2 * c->num_sms + nbx - 1) / nbx));
  chunks = std::max(chunks, (c->M + PREP_MAX_ROWS - 1) / PREP_MAX_ROWS);
  chunks = std::max(chunks, (c->mloc + nbx - 1) / nbx);     // z = 2 needs nbx * chunks >= mloc CTAs
  chunks = std::min(chunks, 65535);
  pp.rows_per_chunk = (c->M + chunks - 1) / chunks;
  CU_TRY(launch_chain(c, k_prepare_q<NT>, dim3(nbx, chunks, 3), dim3(256),
                      (size_t)2 * c->M * sizeof(double), pp));
  c->launches++;
  return OO_OK;
}

int launch_prep(oo_ctx* c, const double* U, const int* done_flag, int kindA, int kindB,
                bool onebody_only = false) {
  switch (c->NT) {
    case 1: return launch_prep_t<1>(c, U, done_flag, kindA, kindB, onebody_only);
    case 2: return launch_prep_t<2>(c, U, done_flag, kindA, kindB, onebody_only);
    case 3: return launch_prep_t<3>(c, U, done_flag, kindA, kindB, onebody_only);
    case 4: return launch_prep_t<4>(c, U, done_flag, kindA, kindB, onebody_only);
  }
  return fail(OO_ERR_INVALID, "unsupported N");
}

template <int NT>
int launch_qc_t(oo_ctx* c, const int* done_flag, bool dense_mirror) {
  constexpr int Np = NT * 8;
  const size_t smem = qc_smem_bytes(NT, c->M, c->mloc);
  if (smem > 200 * 1024) return fail(OO_ERR_UNSUPPORTED, "M too large for k_qcontract smem");
  static size_t attr_smem[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (smem > 48 * 1024 && attr_smem[c->device & 7] < smem) {
    CU_TRY(cudaFuncSetAttribute(k_qcontract<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
    CU_TRY(cudaFuncSetAttribute(k_qcontract<NT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared));
    attr_smem[c->device & 7] = smem;
  }
  QCParams qp;
  qp.Y = c->Y;
  qp.YT = c->YT;
  qp.Upad = c->Upad;
  qp.T3 = c->T3;
  const bool pair = c->pair_sym && !c->generic;
  qp.rowstart = pair ? c->rowstart : nullptr;
  qp.dense_mirror = dense_mirror ? 1 : 0;
  qp.done_flag = done_flag;
  qp.M = c->M;
  qp.t0 = c->t0;
  qp.mloc = c->mloc;
  qp.row0 = pair ? 0 : c->t0;
  qp.nrows = pair ? c->M : c->mloc;
  dim3 grid(qp.nrows, (Np * Np + QC_ECHUNK - 1) / QC_ECHUNK);
  k_qcontract<NT><<<grid, QC_THREADS, smem, c->stream>>>(qp);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return OO_OK;
}

template <int NT>
int launch_tail_t(oo_ctx* c, const double* U, double* out, const int* done_flag, bool fused,
                  int pass, const StepParams* step) {
  TailReduceParams tp;
  memset(&tp, 0, sizeof tp);
  if (fused) fill_peer_comm(c, tp.comm);
  const bool pair = c->pair_sym && !c->generic;
  tp.Aslab = c->Aslab;
  tp.rowstart = pair ? c->rowstart : nullptr;
  tp.U = U;
  tp.B1 = c->B1;
  tp.B12 = c->B12;
  tp.out = out;
  tp.rowE = c->rowE;
  tp.counter = c->counter;
  tp.done_flag = done_flag;
  tp.M = c->M;
  tp.N = c->N;
  tp.t0 = c->t0;
  tp.mloc = c->mloc;
  const bool all_rows = pair || c->generic;
  tp.row0 = all_rows ? 0 : c->t0;
  tp.nrows = all_rows ? c->M : c->mloc;
  tp.mirror_mode = c->generic ? 2 : (pair ? 1 : 0);
  tp.energy_mirror = c->generic ? 0 : 1;
  tp.grad_factor = c->generic ? 1.0 : 4.0;
  tp.accumulate = pass > 0 ? 1 : 0;
  tp.do_step = step != nullptr ? 1 : 0;
  if (step && c->M * c->N <= STEP_SMALL_MN && getenv("OO_NO_SMALL_STEP") == nullptr) tp.do_step = 2;
  if (step) tp.step = *step;
  // V of the optimiser transition in dynamic shared memory (see opt_step_cta) when it fits
  size_t dyn = 0;
  {
    const size_t need = (size_t)c->M * c->N * sizeof(double);
    static const bool off = getenv("OO_NO_STEP_SMEM") != nullptr;
    if (tp.do_step == 1 && !off && need <= (size_t)100 * 1024) {
      static bool attr_set[8] = {false, false, false, false, false, false, false, false};
      if (!attr_set[c->device & 7]) {
        CU_TRY(cudaFuncSetAttribute(k_tail_reduce<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    100 * 1024));
        attr_set[c->device & 7] = true;
      }
      dyn = need;
      tp.step_smem = 1;
    }
  }
  CU_TRY(launch_chain(c, k_tail_reduce<NT>, dim3(tp.nrows), dim3(TAIL_THREADS), dyn, tp));
  c->launches++;
  return OO_OK;
}

int launch_tail(oo_ctx* c, const double* U, double* out, const int* done_flag, bool fused,
                int pass, const StepParams* step) {
  switch (c->NT) {
    case 1: return launch_tail_t<1>(c, U, out, done_flag, fused, pass, step);
    case 2: return launch_tail_t<2>(c, U, out, done_flag, fused, pass, step);
    case 3: return launch_tail_t<3>(c, U, out, done_flag, fused, pass, step);
    case 4: return launch_tail_t<4>(c, U, out, done_flag, fused, pass, step);
  }
  return fail(OO_ERR_INVALID, "unsupported N");
}

// Tiles path: 2-RDM contraction of T3 (k_tail_row), energy, all-reduce, optional optimiser step.
template <int NT>
int launch_tail_row_t(oo_ctx* c, const double* U, double* out, const int* done_flag, bool fused,
                      const StepParams* step) {
  TailParams tp;
  memset(&tp, 0, sizeof tp);
  if (fused) fill_peer_comm(c, tp.comm);
  tp.T3 = c->T3;
  tp.Gp = c->Gp;
  tp.U = U;
  tp.B1 = c->B1;
  tp.B12 = c->B12;
  tp.out = out;
  tp.rowE = c->rowE;
  tp.counter = c->counter;
  tp.done_flag = done_flag;
  tp.M = c->M;
  tp.N = c->N;
  tp.t0 = c->t0;
  tp.mloc = c->mloc;
  tp.row0 = c->pair_sym ? 0 : c->t0;
  tp.nrows = c->pair_sym ? c->M : c->mloc;
  tp.two_body_grad_factor = 4.0;
  tp.accumulate = 0;
  tp.do_step = step != nullptr ? 1 : 0;
  if (step) tp.step = *step;
  constexpr int R = tail_rows(NT), AC = tail_ac(NT);
  dim3 grid((tp.nrows + R - 1) / R, (NT * 8) / AC);
  CU_TRY(launch_chain(c, k_tail_row<NT>, grid, dim3(TAIL_THREADS), 0, tp));
  c->launches++;
  return OO_OK;
}

int launch_tail_row(oo_ctx* c, const double* U, double* out, const int* done_flag, bool fused,
                    const StepParams* step) {
  switch (c->NT) {
    case 1: return launch_tail_row_t<1>(c, U, out, done_flag, fused, step);
    case 2: return launch_tail_row_t<2>(c, U, out, done_flag, fused, step);
    case 3: return launch_tail_row_t<3>(c, U, out, done_flag, fused, step);
    case 4: return launch_tail_row_t<4>(c, U, out, done_flag, fused, step);
  }
  return fail(OO_ERR_INVALID, "unsupported N");
}

int launch_qc(oo_ctx* c, const int* done_flag, bool dense_mirror = false) {
  switch (c->NT) {
    case 1: return launch_qc_t<1>(c, done_flag, dense_mirror);
    case 2: return launch_qc_t<2>(c, done_flag, dense_mirror);
    case 3: return launch_qc_t<3>(c, done_flag, dense_mirror);
    case 4: return launch_qc_t<4>(c, done_flag, dense_mirror);
  }
  return fail(OO_ERR_INVALID, "unsupported N");
}

// One evaluation: (partial) dE/dU and E into `out` (device, M*N+1).
int do_allreduce(oo_ctx* c, double* buf, size_t count);

// reduce = false: this GPU's partial only.  reduce = true: the sum over all GPUs, through the
// all-reduce fused into the tail kernel (peer memory) when attached, else through NCCL.
// step != NULL: the optimiser transition follows the evaluation (oo_optimize) -- inside the tail
// kernel's last CTA when the reduced result is available there, else as a separate k_step launch.
int enqueue_eval(oo_ctx* c, const double* U, double* out, const int* done_flag, bool reduce,
                 const StepParams* step = nullptr) {
  if (!c->have_ints) return fail(OO_ERR_STATE, "oo_set_integrals has not been called");
  if (!c->have_rdms) return fail(OO_ERR_STATE, "oo_set_rdms has not been called");
  const bool tm = c->timing;
  const bool pair = c->pair_sym && !c->generic;
  const bool fused = reduce && c->peer_on && c->world > 1;
  const bool nccl = reduce && !fused && c->world > 1;
  const StepParams* step_in_tail = (step && !nccl && c->step_fusable) ? step : nullptr;
  int rc;
  if (tm) CU_TRY(cudaEventRecord(c->ev[0], c->stream));
  if (c->generic) {
    // no V4 symmetry: one term per index slot.  Slots 0,1 (rows t, q) come from
    // the pass over g, slots 2,3 (rows r, s) from the pass over the pair-transposed tensor.
    if ((rc = launch_prep(c, U, done_flag, 2, 3))) return rc;
    if (tm) CU_TRY(cudaEventRecord(c->ev[1], c->stream));
    if ((rc = launch_k1(c, U, done_flag, false, true))) return rc;
    if (tm) CU_TRY(cudaEventRecord(c->ev[2], c->stream));
    if ((rc = launch_tail(c, U, out, done_flag, false, 0, nullptr))) return rc;
    if ((rc = launch_prep(c, U, done_flag, 4, 5))) return rc;
    if ((rc = launch_k1(c, U, done_flag, true, true))) return rc;
    if ((rc = launch_tail(c, U, out, done_flag, fused, 1, step_in_tail))) return rc;
  } else if (c->tiles_eval) {
    // tiles path (N in 25..32): one-body rows, K1 storing the tiles, q-contraction, T3 x 2-RDM
    if ((rc = launch_prep(c, U, done_flag, 0, -1, true))) return rc;
    if (tm) CU_TRY(cudaEventRecord(c->ev[1], c->stream));
    if ((rc = launch_k1(c, U, done_flag, false, false))) return rc;
    if (tm) CU_TRY(cudaEventRecord(c->ev[2], c->stream));
    if ((rc = launch_qc(c, done_flag))) return rc;
    if (!pair && c->mloc < c->M) {
      const size_t N = (size_t)c->N;
      if (c->t0 > 0) CU_TRY(cudaMemsetAsync(out, 0, (size_t)c->t0 * N * sizeof(double), c->stream));
      const size_t end = (size_t)(c->t0 + c->mloc);
      if (end < (size_t)c->M)
        CU_TRY(cudaMemsetAsync(out + end * N, 0, ((size_t)c->M - end) * N * sizeof(double),
                               c->stream));
    }
    if ((rc = launch_tail_row(c, U, out, done_flag, fused, step_in_tail))) return rc;
  } else {
    if ((rc = launch_prep(c, U, done_flag, 0, pair ? 1 : -1))) return rc;
    if (tm) CU_TRY(cudaEventRecord(c->ev[1], c->stream));
    if ((rc = launch_k1(c, U, done_flag, false, true))) return rc;
    if (tm) CU_TRY(cudaEventRecord(c->ev[2], c->stream));
    if (!pair && c->mloc < c->M) {
      // dense mode writes only the shard's rows; an in-place all-reduce of the previous
      // evaluation may have left full rows elsewhere, so clear them
      const size_t N = (size_t)c->N;
      if (c->t0 > 0) CU_TRY(cudaMemsetAsync(out, 0, (size_t)c->t0 * N * sizeof(double), c->stream));
      const size_t end = (size_t)(c->t0 + c->mloc);
      if (end < (size_t)c->M)
        CU_TRY(cudaMemsetAsync(out + end * N, 0, ((size_t)c->M - end) * N * sizeof(double),
                               c->stream));
    }
    if ((rc = launch_tail(c, U, out, done_flag, fused, 0, step_in_tail))) return rc;
  }
  if (tm) CU_TRY(cudaEventRecord(c->ev[3], c->stream));
  if (nccl && (rc = do_allreduce(c, out, (size_t)c->M * c->N + 1))) return rc;
  if (step && !step_in_tail) {
    // separate transition launch (NCCL all-reduce, OO_NO_STEP_FUSION): V in dynamic shared memory
    const size_t need = (size_t)c->M * c->N * sizeof(double);
    static const bool smem_off = getenv("OO_NO_STEP_SMEM") != nullptr;
    const bool in_smem = !smem_off && need <= (size_t)100 * 1024;
    static bool attr_set[8] = {false, false, false, false, false, false, false, false};
    if (in_smem && !attr_set[c->device & 7]) {
      CU_TRY(cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_set[c->device & 7] = true;
    }
    k_step<<<1, K3_THREADS, in_smem ? need : 0, c->stream>>>(*step, in_smem ? 1 : 0);
    CU_TRY(cudaGetLastError());
    c->launches++;
  }
  return OO_OK;
}

int do_allreduce(oo_ctx* c, double* buf, size_t count) {
  if (!c->comm) return OO_OK;
  int r = g_nccl.all_reduce(buf, buf, count, kNcclFloat64, kNcclSum, c->comm, c->stream);
  if (r != 0) return fail(OO_ERR_NCCL, "ncclAllReduce: %s", g_nccl.get_error_string(r));
  return OO_OK;
}

int alloc_zero(double** p, size_t n) {
  if (*p) return OO_OK;
  CU_TRY(cudaMalloc((void**)p, n * sizeof(double)));
  CU_TRY(cudaMemset(*p, 0, n * sizeof(double)));
  return OO_OK;
}

// Workspaces of the evaluation: 2-RDM layouts (kinds 0,1; 2..5 for the generic path), Q tensors,
// per-slab records.
int ensure_eval_ws(oo_ctx* c, bool generic) {
  const size_t Np = c->Np, Np3 = Np * Np * Np;
  int rc;
  for (int k = generic ? 2 : 0; k < (generic ? 6 : 2); ++k)
    if ((rc = alloc_zero(&c->G2[k], Np3 * Np))) return rc;
  if ((rc = alloc_zero(&c->QA, (size_t)c->M * Np3))) return rc;
  if ((rc = alloc_zero(&c->QB, (size_t)c->mloc * Np3))) return rc;
  if ((rc = alloc_zero(&c->Aslab, (size_t)c->mloc * c->M * 2 * Np))) return rc;
  return OO_OK;
}

// Workspaces of oo_transform (tile mode of K1, q-contraction).
int ensure_transform_ws(oo_ctx* c) {
  const size_t Np2 = (size_t)c->Np * c->Np;
  int rc;
  if ((rc = alloc_zero(&c->Y, (size_t)c->mloc * c->M * Np2))) return rc;
  if ((rc = alloc_zero(&c->YT, (size_t)c->mloc * (c->M / 2 + 1) * Np2))) return rc;
  if ((rc = alloc_zero(&c->Upad, (size_t)c->M * c->Np))) return rc;
  if ((rc = alloc_zero(&c->T3, (size_t)c->M * c->Np * Np2))) return rc;
  return OO_OK;
}

// 2-RDM in the layouts the evaluation needs (G_dev: [N]^4 spatial, spin-summed, weighted).
int prepare_gammas(oo_ctx* c, const double* G_dev) {
  int rc = ensure_eval_ws(c, c->generic);
  if (rc) return rc;
  if (c->tiles_eval && !c->generic) {
    if ((rc = ensure_transform_ws(c))) return rc;
    if ((rc = alloc_zero(&c->Gp, (size_t)c->Np * c->Np * c->Np * c->Np))) return rc;
    k_prepare_gamma<<<c->N * c->Np, 256, 0, c->stream>>>(G_dev, c->Gp, c->N, c->Np, 1);
    CU_TRY(cudaGetLastError());
    c->launches++;
  }
  for (int k = c->generic ? 2 : 0; k < (c->generic ? 6 : 2); ++k) {
    k_prepare_gamma2<<<c->Np * c->Np, 256, 0, c->stream>>>(G_dev, c->G2[k], c->N, c->Np, k);
    CU_TRY(cudaGetLastError());
    c->launches++;
  }
  return OO_OK;
}

// The fused all-reduce gave up waiting for a peer (slow callback, hung rank): every result since
// then is NaN-poisoned on the device; report it from every synchronous entry point.
int check_peer(oo_ctx* c) {
  if (c->peer_on && c->peer_err_host && *c->peer_err_host)
    return fail(OO_ERR_NCCL,
                "fused all-reduce timed out after %.1f s waiting for a peer GPU; results since then "
                "are invalid (NaN).  A host callback or a stalled rank can cause this: raise "
                "OO_PEER_TIMEOUT_MS / oo_set_peer_timeout_ms or use the NCCL all-reduce",
                c->peer_timeout_ns * 1e-9);
  return OO_OK;
}

// Lazily created resources of the pipelined host-buffer evaluation (two slots).
int ensure_slots(oo_ctx* c) {
  if (c->slot_u[0]) return OO_OK;
  const size_t MN = (size_t)c->M * c->N;
  for (int s = 0; s < 2; ++s) {
    CU_TRY(cudaMallocHost((void**)&c->slot_pin_u[s], MN * sizeof(double)));
    CU_TRY(cudaMallocHost((void**)&c->slot_pin_out[s], (MN + 1) * sizeof(double)));
    CU_TRY(cudaMalloc((void**)&c->slot_u[s], MN * sizeof(double)));
    CU_TRY(cudaMalloc((void**)&c->slot_out[s], (MN + 1) * sizeof(double)));
    CU_TRY(cudaMemset(c->slot_out[s], 0, (MN + 1) * sizeof(double)));
    CU_TRY(cudaEventCreateWithFlags(&c->slot_h2d[s], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&c->slot_eval[s], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&c->slot_done[s], cudaEventDisableTiming));
  }
  CU_TRY(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  CU_TRY(cudaDeviceSynchronize());
  return OO_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* oo_last_error(void) { return g_last_error.c_str(); }
const char* oo_version(void) { return "oo_b200 0.2 (sm_100a)"; }

int oo_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return fail(OO_ERR_CUDA, "no CUDA device visible");
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, d) == cudaSuccess && pr.major == 10) ++ok;
  }
  return ok;
}

int oo_create(int device, int M, int N, int t0, int mloc, oo_ctx** out) {
  if (!out) return fail(OO_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (M < 1 || N < 1 || N > M) return fail(OO_ERR_INVALID, "need 1 <= N <= M (got M=%d N=%d)", M, N);
  if (N > K3_NMAX) return fail(OO_ERR_UNSUPPORTED, "N=%d > %d is not supported", N, K3_NMAX);
  if (M % 2) return fail(OO_ERR_UNSUPPORTED, "M must be even (pad the integrals); got %d", M);
  if (t0 < 0 || mloc < 1 || t0 + mloc > M)
    return fail(OO_ERR_INVALID, "bad shard [%d,%d) for M=%d", t0, t0 + mloc, M);
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(OO_ERR_INVALID, "device %d of %d", device, ndev);
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(OO_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is sm_100a only", device,
                prop.major, prop.minor);
  CU_TRY(cudaSetDevice(device));
  oo_ctx* c = new oo_ctx();
  c->device = device;
  c->M = M;
  c->N = N;
  c->NT = (N + 7) / 8;
  c->Np = c->NT * 8;
  c->t0 = t0;
  c->mloc = mloc;
  c->num_sms = prop.multiProcessorCount;
  {
    const char* fj = getenv("OO_FORCE_JACOBI");
    c->force_jacobi = (fj && *fj && *fj != '0') ? 1 : 0;
  }
  c->Mk = (M + K1_KC - 1) / K1_KC * K1_KC;
  c->box_rows = M >= K1_ROWS ? K1_ROWS : (M + 7) / 8 * 8;
  // deepest TMA ring that fits 227 KiB; folding the 8 per-warp partial tiles into 4 or 2
  // buffers is used only when it buys another stage
  c->nstage = 0;
  c->npart = K1_NWARP;
  for (int npart : {8, 4, 2}) {
    for (int ns = 6; ns >= 2; --ns) {
      if (k1_smem_bytes(c->NT, c->Mk, ns, npart) <= (size_t)227 * 1024) {
        if (ns > c->nstage) {
          c->nstage = ns;
          c->npart = npart;
        }
        break;
      }
    }
  }
  if (!c->nstage) {
    delete c;
    return fail(OO_ERR_UNSUPPORTED, "M=%d N=%d does not fit the K1 shared-memory plan", M, N);
  }
  c->k1_smem = k1_smem_bytes(c->NT, c->Mk, c->nstage, c->npart);
  const size_t MN = (size_t)M * N, Np2 = (size_t)c->Np * c->Np;
  auto alloc = [&](double** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(double)); };
  cudaError_t e = cudaSuccess;
  auto A = [&](double** p, size_t n) {
    if (e == cudaSuccess) e = alloc(p, n);
    if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(double));
  };
  // the problem-sized workspaces (Q tensors, Aslab; Y / YT / T3 of oo_transform) are allocated by
  // the first call that needs them: a context that only serves oo_orth / oo_bb_update stays small
  (void)Np2;
  A(&c->Gtmp, (size_t)N * N * N * N);
  A(&c->B1, (size_t)mloc * N);
  A(&c->B12, (size_t)mloc * N);
  A(&c->D, (size_t)N * N);
  A(&c->rowE, (size_t)4 * M);
  A(&c->out, MN + 1);
  A(&c->Ucur, MN);
  A(&c->Uprev, MN);
  A(&c->Gprev, MN);
  {
    // slab tables of the pair-symmetric mode (closed forms: pair_row_count / pair_ith_q)
    std::vector<int> coord, rowstart(mloc, 0);
    coord.reserve((size_t)mloc * (M / 2 + 1));
    for (int tl = 0; tl < mloc; ++tl) {
      rowstart[tl] = (int)coord.size();
      const int t = t0 + tl, cnt = pair_row_count(t, M);
      for (int i = 0; i < cnt; ++i) coord.push_back(tl * M + pair_ith_q(t, i));
    }
    c->nsel = (int)coord.size();
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->slab_coord, coord.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->rowstart, rowstart.size() * sizeof(int));
    if (e == cudaSuccess)
      e = cudaMemcpy(c->slab_coord, coord.data(), coord.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
      e = cudaMemcpy(c->rowstart, rowstart.data(), rowstart.size() * sizeof(int),
                     cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->counter, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(c->counter, 0, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->state, sizeof(OptState));
  if (e == cudaSuccess) e = cudaMemset(c->state, 0, sizeof(OptState));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&c->pin, (MN + 1) * sizeof(double));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&c->pin_state, 2 * sizeof(OptState));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&c->pin_hist, 2 * 8 * sizeof(double));
  // A blocking stream: it orders itself against the legacy default stream, which is where a host
  // framework (torch) produces the input tensors unless told otherwise.
  if (e == cudaSuccess) e = cudaStreamCreate(&c->stream);
  c->own_stream = (e == cudaSuccess && c->stream != nullptr);   // so that a failed create frees it
  {
    const char* nf = getenv("OO_NO_STEP_FUSION");
    c->step_fusable = !(nf && *nf && *nf != '0');
    // N <= 24: the 2-RDM contraction is fused into K1's epilogue; N in 25..32: the epilogue's two
    // dot-product warps would need 2048 FP64 instructions per slab and stall the consumers
    // (measured: K1 1.33 -> 2.10 ms at M=256, N=32), so K1 stores the tiles and
    // k_qcontract / k_tail_row finish the evaluation.  OO_EVAL_PATH=fused|tiles overrides.
    c->tiles_eval = c->NT >= 4;
    if (const char* ep = getenv("OO_EVAL_PATH")) {
      if (!strcmp(ep, "tiles")) c->tiles_eval = true;
      if (!strcmp(ep, "fused")) c->tiles_eval = false;
    }
  }
  for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i)
    e = cudaEventCreateWithFlags(&c->poll_ev[i], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    int rc = fail(OO_ERR_CUDA, "allocation failed: %s", cudaGetErrorString(e));
    oo_destroy(c);
    return rc;
  }
  *out = c;
  return OO_OK;
}

int oo_destroy(oo_ctx* c) {
  if (!c) return OO_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  // a graph that captured NCCL work must go before the communicator it refers to
  if (c->chunk_graph) {
    cudaGraphExecDestroy(c->chunk_graph);
    c->chunk_graph = nullptr;
  }
  if (c->comm && g_nccl.comm_destroy) g_nccl.comm_destroy(c->comm);
  for (int r = 0; r < PEER_MAX; ++r)
    if (c->peer_map[r] && c->peer_map[r] != c->peer_base) cudaIpcCloseMemHandle(c->peer_map[r]);
  for (int s2 = 0; s2 < 6; ++s2)
    if (c->G2[s2]) cudaFree(c->G2[s2]);
  if (c->Gp) cudaFree(c->Gp);
  if (c->QA) cudaFree(c->QA);
  if (c->QB) cudaFree(c->QB);
  if (c->Aslab) cudaFree(c->Aslab);
  if (c->peer_base) cudaFree(c->peer_base);
  if (c->peer_err) cudaFree(c->peer_err);
  if (c->peer_err_host) cudaFreeHost((void*)c->peer_err_host);
  for (int s2 = 0; s2 < 2; ++s2) {
    if (c->slot_pin_u[s2]) cudaFreeHost(c->slot_pin_u[s2]);
    if (c->slot_pin_out[s2]) cudaFreeHost(c->slot_pin_out[s2]);
    if (c->slot_u[s2]) cudaFree(c->slot_u[s2]);
    if (c->slot_out[s2]) cudaFree(c->slot_out[s2]);
    if (c->slot_h2d[s2]) cudaEventDestroy(c->slot_h2d[s2]);
    if (c->slot_eval[s2]) cudaEventDestroy(c->slot_eval[s2]);
    if (c->slot_done[s2]) cudaEventDestroy(c->slot_done[s2]);
  }
  if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  if (c->peer_seq_dev) cudaFree(c->peer_seq_dev);
  double* bufs[] = {c->Y,   c->T3,   c->D,     c->rowE,
                    c->out, c->Ucur, c->Uprev, c->Gprev, c->E_hist,
                    c->YT,  c->Upad, c->B1,    c->B12,   c->Gtmp};
  for (double* b : bufs)
    if (b) cudaFree(b);
  if (c->counter) cudaFree(c->counter);
  if (c->slab_coord) cudaFree(c->slab_coord);
  if (c->rowstart) cudaFree(c->rowstart);
  if (c->state) cudaFree(c->state);
  if (c->pin) cudaFreeHost(c->pin);
  if (c->pin_state) cudaFreeHost(c->pin_state);
  if (c->pin_hist) cudaFreeHost(c->pin_hist);
  for (auto& ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->poll_ev)
    if (ev) cudaEventDestroy(ev);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return OO_OK;
}

int oo_set_stream(oo_ctx* c, void* cuda_stream) {
  if (!c) return fail(OO_ERR_INVALID, "ctx is NULL");
  CU_TRY(cudaSetDevice(c->device));
  if (c->own_stream && c->stream) {
    CU_TRY(cudaStreamSynchronize(c->stream));
    CU_TRY(cudaStreamDestroy(c->stream));
    c->own_stream = false;
  }
  if (cuda_stream) {
    c->stream = (cudaStream_t)cuda_stream;
  } else {
    CU_TRY(cudaStreamCreate(&c->stream));
    c->own_stream = true;
  }
  return OO_OK;
}

int oo_synchronize(oo_ctx* c) {
  if (!c) return fail(OO_ERR_INVALID, "ctx is NULL");
  CU_TRY(cudaStreamSynchronize(c->stream));
  return check_peer(c);
}

int oo_set_integrals(oo_ctx* c, const double* h_dev, const double* g_dev, unsigned flags) {
  if (!c || !h_dev || !g_dev) return fail(OO_ERR_INVALID, "NULL argument");
  if (((uintptr_t)g_dev & 15) != 0) return fail(OO_ERR_INVALID, "g must be 16-byte aligned");
  if (!(flags & OO_G_V4_SYMMETRIC))
    return fail(OO_ERR_UNSUPPORTED,
                "two-body tensors without V4 symmetry (g[pqrs]=g[qpsr]=g[rspq]) need the "
                "pair-transposed copy as well: use oo_set_integrals_generic");
  CU_TRY(cudaSetDevice(c->device));
  const bool packed = (flags & OO_G_PAIR_PACKED) != 0;
  int rc = build_tmap(c, g_dev, &c->tmap, packed ? (size_t)c->nsel : 0);
  if (rc) return rc;
  if ((rc = ensure_eval_ws(c, false))) return rc;
  if (c->generic) c->have_rdms = false;   // the 2-RDM layouts of the symmetric path are not prepared
  c->h = h_dev;
  c->g = g_dev;
  c->gflags = flags;
  c->generic = false;
  c->packed = packed;
  if (packed && !c->pair_sym) {
    // packed storage only holds what the pair-symmetric mode streams
    if ((rc = oo_set_pair_symmetry(c, 1))) return rc;
  }
  c->have_ints = true;
  return OO_OK;
}

int oo_pair_slab_list(int M, int t0, int mloc, int* tq_host, int capacity) {
  if (M < 1 || t0 < 0 || mloc < 1 || t0 + mloc > M) return fail(OO_ERR_INVALID, "bad shard");
  long n = 0;
  for (int t = t0; t < t0 + mloc; ++t) {
    const int cnt = pair_row_count(t, M);
    for (int i = 0; i < cnt; ++i, ++n) {
      if (tq_host && n < capacity) {
        tq_host[2 * n] = t;
        tq_host[2 * n + 1] = pair_ith_q(t, i);
      }
    }
  }
  if (tq_host && n > capacity) return fail(OO_ERR_INVALID, "tq_host holds %d of %ld slabs", capacity, n);
  return (int)n;
}

int oo_pack_pair_slabs(oo_ctx* c, const double* g_dense_dev, double* g_packed_dev) {
  if (!c || !g_dense_dev || !g_packed_dev) return fail(OO_ERR_INVALID, "NULL argument");
  if ((((uintptr_t)g_dense_dev) | ((uintptr_t)g_packed_dev)) & 15)
    return fail(OO_ERR_INVALID, "g must be 16-byte aligned");
  CU_TRY(cudaSetDevice(c->device));
  const int chunks = std::max(1, std::min(64, (int)(((size_t)c->M * c->M / 2 + 255) / 256)));
  k_pack_slabs<<<dim3(c->nsel, chunks), 256, 0, c->stream>>>(g_dense_dev, g_packed_dev,
                                                            c->slab_coord, c->M);
  CU_TRY(cudaGetLastError());
  return OO_OK;
}

int oo_set_integrals_generic(oo_ctx* c, const double* h_dev, const double* g_dev,
                             const double* g_pair_transposed_dev) {
  if (!c || !h_dev || !g_dev || !g_pair_transposed_dev) return fail(OO_ERR_INVALID, "NULL argument");
  if ((((uintptr_t)g_dev) | ((uintptr_t)g_pair_transposed_dev)) & 15)
    return fail(OO_ERR_INVALID, "g must be 16-byte aligned");
  CU_TRY(cudaSetDevice(c->device));
  c->h = h_dev;
  c->g = g_dev;
  c->g2 = g_pair_transposed_dev;
  c->gflags = 0;
  c->packed = false;
  int rc = build_tmap(c, c->g, &c->tmap);
  if (rc) return rc;
  if ((rc = build_tmap(c, c->g2, &c->tmap2))) return rc;
  if ((rc = ensure_eval_ws(c, true))) return rc;
  if (!c->generic) c->have_rdms = false;   // the 2-RDM must be re-ingested in the four slot layouts
  c->generic = true;
  c->have_ints = true;
  return OO_OK;
}

int oo_check_v4_symmetry(int device, const double* g_dev, int M, double* out_host) {
  if (!g_dev || !out_host || M < 1) return fail(OO_ERR_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(device));
  unsigned long long* d = nullptr;
  CU_TRY(cudaMalloc((void**)&d, 2 * sizeof(unsigned long long)));
  CU_TRY(cudaMemset(d, 0, 2 * sizeof(unsigned long long)));
  {
    const int T = (M + 31) / 32;
    dim3 grid((unsigned)((long)M * (M + 1) / 2), T, T), block(32, 8);
    k_v4_symmetry_tiles<0><<<grid, block>>>(g_dev, M, d);
    k_v4_symmetry_tiles<1><<<grid, block>>>(g_dev, M, d);
  }
  cudaError_t e = cudaGetLastError();
  unsigned long long hbits[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpy(hbits, d, sizeof hbits, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(OO_ERR_CUDA, "symmetry check: %s", cudaGetErrorString(e));
  memcpy(out_host, hbits, sizeof hbits);
  return OO_OK;
}

int oo_set_rdms(oo_ctx* c, const double* D_dev, const double* G_dev) {
  if (!c || !D_dev || !G_dev) return fail(OO_ERR_INVALID, "NULL argument");
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaMemcpyAsync(c->D, D_dev, (size_t)c->N * c->N * sizeof(double),
                         cudaMemcpyDeviceToDevice, c->stream));
  int rc = prepare_gammas(c, G_dev);
  if (rc) return rc;
  c->have_rdms = true;
  return OO_OK;
}

int oo_ingest_spin_g_rows(int device, const double* g_spin_dev, int M, double rtol, int t0,
                          int mloc, int Mpad, double* g_out_dev, unsigned* block_mask,
                          double* stats_host) {
  if (!g_spin_dev || !g_out_dev || !block_mask || M < 1 || t0 < 0 || mloc < 1 || t0 + mloc > M ||
      Mpad < M)
    return fail(OO_ERR_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(device));
  unsigned long long* d = nullptr;
  CU_TRY(cudaMalloc((void**)&d, 32 * sizeof(unsigned long long)));
  CU_TRY(cudaMemset(d, 0, 32 * sizeof(unsigned long long)));
  k_spin_block_maxabs<<<148 * 4, 256>>>(g_spin_dev, M, d);
  unsigned long long bits[32];
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(bits, d, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(OO_ERR_CUDA, "spin-block scan: %s", cudaGetErrorString(e));
  }
  double bmax[16], gmax = 0.0;
  memcpy(bmax, bits, sizeof bmax);
  for (int b = 0; b < 16; ++b) gmax = std::max(gmax, bmax[b]);
  unsigned mask = 0;
  int ref = -1;
  for (int b = 0; b < 16; ++b)
    if (bmax[b] > rtol * (gmax > 0 ? gmax : 1.0)) {
      mask |= 1u << b;
      if (ref < 0) ref = b;
    }
  if (ref < 0) ref = 0;  // all-zero tensor: take the alpha-alpha-alpha-alpha block
  k_spin_block_extract<<<148 * 4, 256>>>(g_spin_dev, M, ref, mask, t0, mloc, Mpad, g_out_dev, d + 16);
  e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpy(bits, d + 16, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(OO_ERR_CUDA, "spin-block extract: %s", cudaGetErrorString(e));
  double dmax[16], dev_max = 0.0;
  memcpy(dmax, bits, sizeof dmax);
  for (int b = 0; b < 16; ++b) dev_max = std::max(dev_max, dmax[b]);
  *block_mask = mask;
  if (stats_host) {
    stats_host[0] = gmax;
    stats_host[1] = dev_max;
  }
  if (dev_max > rtol * (gmax > 0 ? gmax : 1.0))
    return fail(OO_ERR_UNSUPPORTED,
                "unrestricted two-body integrals: non-zero spin blocks differ (max deviation %.3e, "
                "max |g| %.3e)", dev_max, gmax);
  return OO_OK;
}

int oo_ingest_spin_g(int device, const double* g_spin_dev, int M, double rtol,
                     double* g_sp_out_dev, unsigned* block_mask, double* stats_host) {
  return oo_ingest_spin_g_rows(device, g_spin_dev, M, rtol, 0, M, M, g_sp_out_dev, block_mask,
                               stats_host);
}

int oo_set_rdms_spin(oo_ctx* c, const double* const* D_spin_dev, const double* const* G_spin_dev,
                     const double* weights_host, int nstates, unsigned block_mask) {
  if (!c || !D_spin_dev || !G_spin_dev || nstates < 1)
    return fail(OO_ERR_INVALID, "bad argument");
  if (nstates > 8) return fail(OO_ERR_UNSUPPORTED, "at most 8 states per call (got %d)", nstates);
  CU_TRY(cudaSetDevice(c->device));
  RdmSpinParams rp;
  memset(&rp, 0, sizeof rp);
  for (int n = 0; n < nstates; ++n) {
    if (!D_spin_dev[n] || !G_spin_dev[n]) return fail(OO_ERR_INVALID, "NULL RDM pointer");
    rp.D[n] = D_spin_dev[n];
    rp.G[n] = G_spin_dev[n];
    rp.w[n] = weights_host ? weights_host[n] : 1.0;
  }
  rp.nstates = nstates;
  rp.N = c->N;
  rp.mask = block_mask;
  const size_t N4 = (size_t)c->N * c->N * c->N * c->N;
  const int grid = (int)std::min<size_t>((N4 + 255) / 256, 148 * 8);
  k_rdm_spin_sum<<<grid, 256, 0, c->stream>>>(rp, c->D, c->Gtmp);
  CU_TRY(cudaGetLastError());
  c->launches++;
  int rc = prepare_gammas(c, c->Gtmp);
  if (rc) return rc;
  c->have_rdms = true;
  return OO_OK;
}

int oo_energy_grad(oo_ctx* c, const double* U_dev, double* out_dev) {
  if (!c || !U_dev) return fail(OO_ERR_INVALID, "NULL argument");
  CU_TRY(cudaSetDevice(c->device));
  return enqueue_eval(c, U_dev, out_dev ? out_dev : c->out, nullptr, false);
}

int oo_energy_grad_allreduce(oo_ctx* c, const double* U_dev, double* out_dev) {
  if (!c || !U_dev) return fail(OO_ERR_INVALID, "NULL argument");
  CU_TRY(cudaSetDevice(c->device));
  return enqueue_eval(c, U_dev, out_dev ? out_dev : c->out, nullptr, true);
}

int oo_eval_submit(oo_ctx* c, const double* U_host, int slot) {
  if (!c || !U_host || slot < 0 || slot > 1) return fail(OO_ERR_INVALID, "bad argument");
  if (!c->have_ints || !c->have_rdms) return fail(OO_ERR_STATE, "integrals / RDMs not set");
  CU_TRY(cudaSetDevice(c->device));
  int rc = ensure_slots(c);
  if (rc) return rc;
  if (c->slot_busy[slot]) return fail(OO_ERR_STATE, "slot %d still holds an unread result", slot);
  const size_t MN = (size_t)c->M * c->N;
  memcpy(c->slot_pin_u[slot], U_host, MN * sizeof(double));
  // H2D on its own stream (overlaps the evaluation that is running), evaluation on the context
  // stream, D2H on a third stream (overlaps the next evaluation)
  CU_TRY(cudaMemcpyAsync(c->slot_u[slot], c->slot_pin_u[slot], MN * sizeof(double),
                         cudaMemcpyHostToDevice, c->h2d_stream));
  CU_TRY(cudaEventRecord(c->slot_h2d[slot], c->h2d_stream));
  CU_TRY(cudaStreamWaitEvent(c->stream, c->slot_h2d[slot], 0));
  if ((rc = enqueue_eval(c, c->slot_u[slot], c->slot_out[slot], nullptr, true))) return rc;
  CU_TRY(cudaEventRecord(c->slot_eval[slot], c->stream));
  CU_TRY(cudaStreamWaitEvent(c->d2h_stream, c->slot_eval[slot], 0));
  CU_TRY(cudaMemcpyAsync(c->slot_pin_out[slot], c->slot_out[slot], (MN + 1) * sizeof(double),
                         cudaMemcpyDeviceToHost, c->d2h_stream));
  CU_TRY(cudaEventRecord(c->slot_done[slot], c->d2h_stream));
  c->slot_busy[slot] = true;
  return OO_OK;
}

int oo_eval_wait(oo_ctx* c, int slot, double* E_host, double* grad_host) {
  if (!c || slot < 0 || slot > 1) return fail(OO_ERR_INVALID, "bad argument");
  if (!c->slot_busy[slot]) return fail(OO_ERR_STATE, "nothing was submitted to slot %d", slot);
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaEventSynchronize(c->slot_done[slot]));
  c->slot_busy[slot] = false;
  const size_t MN = (size_t)c->M * c->N;
  if (E_host) *E_host = c->slot_pin_out[slot][MN];
  if (grad_host) memcpy(grad_host, c->slot_pin_out[slot], MN * sizeof(double));
  return check_peer(c);
}

int oo_energy_grad_host(oo_ctx* c, const double* U_host, double* E_host, double* grad_host) {
  if (!c || !U_host || !E_host) return fail(OO_ERR_INVALID, "NULL argument");
  int rc = oo_eval_submit(c, U_host, 0);
  if (rc) return rc;
  return oo_eval_wait(c, 0, E_host, grad_host);
}

int oo_transform(oo_ctx* c, const double* U_dev, double* h_rot_dev, double* g_rot_dev) {
  if (!c || !U_dev) return fail(OO_ERR_INVALID, "NULL argument");
  if (!c->have_ints) return fail(OO_ERR_STATE, "oo_set_integrals has not been called");
  CU_TRY(cudaSetDevice(c->device));
  int rc;
  if (g_rot_dev) {
    if ((rc = ensure_transform_ws(c))) return rc;
    if ((rc = launch_k1(c, U_dev, nullptr, false, false))) return rc;
    if ((rc = launch_qc(c, nullptr))) return rc;
    k_rotate_g<<<c->N * c->N, 256, 0, c->stream>>>(c->T3, U_dev, g_rot_dev, c->N, c->Np,
                                                   (c->pair_sym && !c->generic) ? 0 : c->t0,
                                                   (c->pair_sym && !c->generic) ? c->M : c->mloc);
    CU_TRY(cudaGetLastError());
    c->launches++;
  }
  if (h_rot_dev) {
    k_rotate_h<<<c->N * c->N, 128, 0, c->stream>>>(c->h, U_dev, h_rot_dev, c->M, c->N, c->t0,
                                                   c->mloc);
    CU_TRY(cudaGetLastError());
    c->launches++;
  }
  return OO_OK;
}

int oo_orth(oo_ctx* c, const double* V_dev, double* U_out_dev) {
  if (!c || !V_dev || !U_out_dev) return fail(OO_ERR_INVALID, "NULL argument");
  CU_TRY(cudaSetDevice(c->device));
  k_orth<<<1, K3_THREADS, 0, c->stream>>>(V_dev, U_out_dev, c->M, c->N, c->force_jacobi);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return OO_OK;
}

int oo_bb_update(oo_ctx* c, int iteration, const double* U_cur_dev, const double* U_prev_dev,
                 const double* G_cur_dev, const double* G_prev_dev, double* alpha_io_dev,
                 double* U_new_dev) {
  if (!c || !U_cur_dev || !G_cur_dev || !alpha_io_dev || !U_new_dev)
    return fail(OO_ERR_INVALID, "NULL argument");
  if (iteration >= 1 && (!U_prev_dev || !G_prev_dev))
    return fail(OO_ERR_INVALID, "previous iterates are required for iteration >= 1");
  CU_TRY(cudaSetDevice(c->device));
  BBParams p;
  p.Ucur = U_cur_dev;
  p.Uprev = U_prev_dev;
  p.Gcur = G_cur_dev;
  p.Gprev = G_prev_dev;
  p.Unew = U_new_dev;
  p.alpha_io = alpha_io_dev;
  p.iteration = iteration;
  p.M = c->M;
  p.N = c->N;
  p.force_jacobi = c->force_jacobi;
  k_bb_update<<<1, K3_THREADS, 0, c->stream>>>(p);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return OO_OK;
}

int oo_optimize(oo_ctx* c, double* U_io_host, double bb0, double tol, int maxiter, double decay,
                double* E_hist_host, int hist_cap, int* n_iter, double* E_final,
                double* bb_final) {
  if (!c || !U_io_host) return fail(OO_ERR_INVALID, "NULL argument");
  if (!c->have_ints || !c->have_rdms) return fail(OO_ERR_STATE, "integrals / RDMs not set");
  CU_TRY(cudaSetDevice(c->device));
  const size_t MN = (size_t)c->M * c->N;
  const int need_hist = std::max(maxiter, 0) + 8;
  if (c->hist_cap < need_hist) {
    if (c->E_hist) CU_TRY(cudaFree(c->E_hist));
    c->E_hist = nullptr;
    CU_TRY(cudaMalloc((void**)&c->E_hist, (size_t)need_hist * sizeof(double)));
    c->hist_cap = need_hist;
  }
  CU_TRY(cudaMemsetAsync(c->E_hist, 0, (size_t)c->hist_cap * sizeof(double), c->stream));
  memcpy(c->pin, U_io_host, MN * sizeof(double));
  CU_TRY(cudaMemcpyAsync(c->Ucur, c->pin, MN * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  OptState init;
  memset(&init, 0, sizeof init);
  init.alpha = bb0;
  init.S[1] = 1.5 * tol;  // St_array = [None, 1.5*tol]  (pupo.py:178)
  init.tol = tol;
  init.decay = decay;
  init.maxiter = maxiter;
  c->pin_state[0] = init;
  CU_TRY(cudaMemcpyAsync(c->state, &c->pin_state[0], sizeof(OptState), cudaMemcpyHostToDevice,
                         c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));  // pin_state[0] is reused below

  StepParams sp;
  sp.st = c->state;
  sp.Ucur = c->Ucur;
  sp.Uprev = c->Uprev;
  sp.gE = c->out;
  sp.Gprev = c->Gprev;
  sp.E_hist = c->E_hist;
  sp.M = c->M;
  sp.N = c->N;
  sp.hist_cap = c->hist_cap;
  sp.force_jacobi = c->force_jacobi;
  const int* done_flag = &c->state->done;

  // Transitions are enqueued in chunks; the stop flag of chunk i is inspected while chunk i+1
  // is already queued, so the device never idles on the host.  Transitions after the stop are
  // no-ops (every kernel tests the flag first).
  const int chunk = 4;
  const long max_transitions = (long)std::max(maxiter, 3) + 4;
  c->stop_requested = 0;
  long enq = 0;
  int slot = 0;
  bool pending[2] = {false, false};
  bool done = false;
  auto chunk_body = [&]() -> int {
    int rc;
    for (int i = 0; i < chunk; ++i)
      if ((rc = enqueue_eval(c, c->Ucur, c->out, done_flag, true, &sp))) return rc;
    return OO_OK;
  };
  // The chunk (4 x [k_prepare_q, K1, k_tail_reduce(+all-reduce, +step)]) is captured once into
  // a CUDA graph and replayed: the inner loop of small problems is launch-bound.  The first chunk
  // runs un-captured (it also performs the one-time function-attribute calls).
  const bool use_graph = !c->timing && getenv("OO_NO_GRAPH") == nullptr;
  const void* key[10] = {c->E_hist, (const void*)(uintptr_t)c->hist_cap, c->stream, c->g, c->g2,
                         c->h, c->comm, (const void*)(uintptr_t)(c->pair_sym + 2 * c->generic +
                                                                 4 * c->peer_on + 8 * c->packed),
                         c->QA, (const void*)(uintptr_t)(c->world + 16 * c->step_fusable)};
  bool first_chunk = true;
  long chunk_start[2] = {0, 0};
  double prev_f = 0.0;          // f(U_{k-1}) carried across chunks for the callback replay
  static_assert(4 <= 8, "pin_hist holds up to 8 energies per slot");
  // Callbacks are delivered while the device keeps running: at every flag poll the energies of
  // the chunk that just finished are handed out with the reference's arguments
  // (k <= 2: (k, f(U_k)); loop iterations k >= 3: (k, f(U_{k-1})), pupo.py:313).
  auto deliver_callbacks = [&](int s) {
    if (!c->cb) return;
    const OptState& stt = c->pin_state[s];
    const long n_exec = stt.done ? stt.k_final : stt.k;   // transitions whose body ran
    const double* hst = c->pin_hist + s * 8;
    for (long k = chunk_start[s]; k < chunk_start[s] + chunk && k < n_exec; ++k) {
      const double fk = hst[k - chunk_start[s]];
      c->cb((int)k, k <= 2 ? fk : prev_f, c->cb_user);
      prev_f = fk;
    }
  };
  auto enqueue_chunk = [&](int s) -> int {
    int rc;
    if (!use_graph || first_chunk) {
      if ((rc = chunk_body())) return rc;
      first_chunk = false;
    } else {
      if (c->chunk_graph && memcmp(key, c->graph_key, sizeof key) != 0) {
        cudaGraphExecDestroy(c->chunk_graph);
        c->chunk_graph = nullptr;
      }
      if (!c->chunk_graph) {
        const long long before = c->launches;
        CU_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        rc = chunk_body();
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        if (rc) {
          if (graph) cudaGraphDestroy(graph);
          return rc;
        }
        if (ce != cudaSuccess)
          return fail(OO_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&c->chunk_graph, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess)
          return fail(OO_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
        c->graph_kernels = c->launches - before;
        c->launches = before;
        memcpy(c->graph_key, key, sizeof key);
      }
      CU_TRY(cudaGraphLaunch(c->chunk_graph, c->stream));
      c->launches += c->graph_kernels;
    }
    chunk_start[s] = enq;
    if (c->cb && enq + chunk <= c->hist_cap)
      CU_TRY(cudaMemcpyAsync(c->pin_hist + s * 8, c->E_hist + enq, chunk * sizeof(double),
                             cudaMemcpyDeviceToHost, c->stream));
    enq += chunk;
    CU_TRY(cudaMemcpyAsync(&c->pin_state[s], c->state, sizeof(OptState), cudaMemcpyDeviceToHost,
                           c->stream));
    CU_TRY(cudaEventRecord(c->poll_ev[s], c->stream));
    pending[s] = true;
    return OO_OK;
  };
  int rc = enqueue_chunk(slot);
  if (rc) return rc;
  while (!done) {
    const int other = slot ^ 1;
    if (enq < max_transitions) {
      if ((rc = enqueue_chunk(other))) return rc;
    }
    CU_TRY(cudaEventSynchronize(c->poll_ev[slot]));
    pending[slot] = false;
    deliver_callbacks(slot);
    if (c->stop_requested == 1) {
      // a callback asked for an early stop: raise the device-side flag behind the queued work
      k_force_stop<<<1, 1, 0, c->stream>>>(c->state);
      CU_TRY(cudaGetLastError());
      c->stop_requested = 2;
    }
    if (c->pin_state[slot].done) {
      done = true;
    } else if (!pending[other]) {
      if (c->stop_requested) {                 // the forced stop is queued behind the last chunk
        CU_TRY(cudaMemcpyAsync(&c->pin_state[other], c->state, sizeof(OptState),
                               cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        if (c->pin_state[other].done) break;
      }
      return fail(OO_ERR_NUMERIC, "optimiser did not stop within %ld transitions", enq);
    }
    slot = other;
  }
  CU_TRY(cudaMemcpyAsync(c->pin, c->Ucur, MN * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemcpyAsync(&c->pin_state[0], c->state, sizeof(OptState), cudaMemcpyDeviceToHost,
                         c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  const OptState fin = c->pin_state[0];
  memcpy(U_io_host, c->pin, MN * sizeof(double));
  if (E_hist_host && hist_cap > 0) {
    const int n = std::min(hist_cap, c->hist_cap);
    CU_TRY(cudaMemcpy(E_hist_host, c->E_hist, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (n_iter) *n_iter = fin.k_final;
  if (E_final) *E_final = fin.E_final;
  if (bb_final) *bb_final = fin.alpha;
  c->last_ns_iters = fin.ns_iters;
  c->last_jacobi_calls = fin.jacobi_calls;
  {
    int prc = check_peer(c);
    if (prc) return prc;
  }
  if (fin.nan_flag) return fail(OO_ERR_NUMERIC, "non-finite value met during the optimisation");
  return OO_OK;
}

int oo_request_stop(oo_ctx* c) {
  if (!c) return fail(OO_ERR_INVALID, "ctx is NULL");
  c->stop_requested = 1;
  return OO_OK;
}

int oo_set_peer_timeout_ms(oo_ctx* c, double milliseconds) {
  if (!c || !(milliseconds > 0)) return fail(OO_ERR_INVALID, "bad argument");
  c->peer_timeout_ns = (unsigned long long)(milliseconds * 1e6);
  return OO_OK;
}

int oo_set_callback(oo_ctx* c, oo_callback_t cb, void* user) {
  if (!c) return fail(OO_ERR_INVALID, "ctx is NULL");
  c->cb = cb;
  c->cb_user = user;
  return OO_OK;
}

int oo_nccl_unique_id(void* id128_host) {
  if (!id128_host) return fail(OO_ERR_INVALID, "NULL argument");
  int rc = load_nccl();
  if (rc) return rc;
  NcclUniqueId128 id;
  int r = g_nccl.get_unique_id(&id);
  if (r != 0) return fail(OO_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.get_error_string(r));
  memcpy(id128_host, &id, sizeof id);
  return OO_OK;
}

int oo_comm_init(oo_ctx* c, const void* id128_host, int rank, int world) {
  if (!c || !id128_host || rank < 0 || rank >= world)
    return fail(OO_ERR_INVALID, "bad communicator arguments");
  int rc = load_nccl();
  if (rc) return rc;
  CU_TRY(cudaSetDevice(c->device));
  NcclUniqueId128 id;
  memcpy(&id, id128_host, sizeof id);
  void* comm = nullptr;
  int r = g_nccl.comm_init_rank(&comm, world, id, rank);
  if (r != 0) return fail(OO_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.get_error_string(r));
  c->comm = comm;
  c->rank = rank;
  c->world = world;
  return OO_OK;
}

int oo_peer_export(oo_ctx* c, void* handle64_host) {
  if (!c || !handle64_host) return fail(OO_ERR_INVALID, "NULL argument");
  CU_TRY(cudaSetDevice(c->device));
  if (!c->peer_base) {
    c->peer_stride = (c->M * c->N + 2) & ~1;
    const size_t bytes = 256 + (size_t)2 * PEER_MAX * c->peer_stride * sizeof(double);
    CU_TRY(cudaMalloc(&c->peer_base, bytes));
    CU_TRY(cudaMemset(c->peer_base, 0, bytes));
    CU_TRY(cudaMalloc((void**)&c->peer_err, sizeof(int)));
    CU_TRY(cudaMemset(c->peer_err, 0, sizeof(int)));
    CU_TRY(cudaHostAlloc((void**)&c->peer_err_host, sizeof(int), cudaHostAllocMapped));
    *c->peer_err_host = 0;
    CU_TRY(cudaHostGetDevicePointer((void**)&c->peer_err_host_dev, (void*)c->peer_err_host, 0));
    if (const char* t = getenv("OO_PEER_TIMEOUT_MS")) {
      const double ms = atof(t);
      if (ms > 0) c->peer_timeout_ns = (unsigned long long)(ms * 1e6);
    }
    CU_TRY(cudaMalloc((void**)&c->peer_seq_dev, sizeof(unsigned long long)));
    CU_TRY(cudaMemset(c->peer_seq_dev, 0, sizeof(unsigned long long)));
    CU_TRY(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  CU_TRY(cudaIpcGetMemHandle(&h, c->peer_base));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64_host, &h, 64);
  return OO_OK;
}

int oo_peer_attach(oo_ctx* c, const void* handles_host, int rank, int world) {
  if (!c || !handles_host || rank < 0 || rank >= world)
    return fail(OO_ERR_INVALID, "bad peer arguments");
  if (world > PEER_MAX) return fail(OO_ERR_UNSUPPORTED, "at most %d peers", PEER_MAX);
  if (!c->peer_base) return fail(OO_ERR_STATE, "call oo_peer_export first");
  CU_TRY(cudaSetDevice(c->device));
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      c->peer_map[r] = c->peer_base;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles_host + 64 * r, 64);
    void* p = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_map[r] = p;
  }
  c->rank = rank;
  c->world = world;
  c->peer_on = true;
  return OO_OK;
}

int oo_peer_status(oo_ctx* c) {
  if (!c) return fail(OO_ERR_INVALID, "ctx is NULL");
  if (!c->peer_on) return 0;
  int err = 0;
  CU_TRY(cudaMemcpy(&err, c->peer_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (err || check_peer(c)) return fail(OO_ERR_NCCL, "fused all-reduce timed out waiting for a peer");
  return 1;
}

int oo_allreduce(oo_ctx* c, double* buf_dev, size_t count) {
  if (!c || !buf_dev) return fail(OO_ERR_INVALID, "NULL argument");
  if (!c->comm) return fail(OO_ERR_STATE, "no communicator attached (oo_comm_init)");
  CU_TRY(cudaSetDevice(c->device));
  return do_allreduce(c, buf_dev, count);
}

int oo_set_pair_symmetry(oo_ctx* c, int enable) {
  if (!c) return fail(OO_ERR_INVALID, "ctx is NULL");
  if (!enable && c->packed)
    return fail(OO_ERR_STATE, "pair-packed integrals hold only the pair-symmetric slab set");
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaStreamSynchronize(c->stream));
  c->pair_sym = enable != 0;
  // rows outside the shard are only written in pair-symmetric mode: start from zero again
  CU_TRY(cudaMemsetAsync(c->out, 0, ((size_t)c->M * c->N + 1) * sizeof(double), c->stream));
  return OO_OK;
}

int oo_streamed_slabs(oo_ctx* c) {
  if (!c) return fail(OO_ERR_INVALID, "ctx is NULL");
  if (c->generic) return 2 * c->mloc * c->M;
  return c->pair_sym ? c->nsel : c->mloc * c->M;
}

int oo_set_timing(oo_ctx* c, int enable) {
  if (!c) return fail(OO_ERR_INVALID, "ctx is NULL");
  c->timing = enable != 0;
  return OO_OK;
}

int oo_last_timing(oo_ctx* c, float* ms5_host) {
  if (!c || !ms5_host) return fail(OO_ERR_INVALID, "NULL argument");
  if (!c->timing) return fail(OO_ERR_STATE, "timing is not enabled");
  CU_TRY(cudaEventSynchronize(c->ev[3]));
  CU_TRY(cudaEventElapsedTime(&ms5_host[0], c->ev[1], c->ev[2]));   // K1
  CU_TRY(cudaEventElapsedTime(&ms5_host[1], c->ev[0], c->ev[1]));   // k_prepare_q
  CU_TRY(cudaEventElapsedTime(&ms5_host[2], c->ev[2], c->ev[3]));   // k_tail_reduce
  ms5_host[3] = 0.f;
  CU_TRY(cudaEventElapsedTime(&ms5_host[4], c->ev[0], c->ev[3]));
  return OO_OK;
}

long long oo_launch_count(oo_ctx* c) { return c ? c->launches : 0; }

int oo_retraction_stats(oo_ctx* c, int* newton_schulz_iterations, int* jacobi_fallbacks) {
  if (!c) return fail(OO_ERR_INVALID, "ctx is NULL");
  if (newton_schulz_iterations) *newton_schulz_iterations = c->last_ns_iters;
  if (jacobi_fallbacks) *jacobi_fallbacks = c->last_jacobi_calls;
  return OO_OK;
}

int oo_measure_peaks(int device, size_t bytes, double* out_host) {
  if (!out_host) return fail(OO_ERR_INVALID, "NULL argument");
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  const int sms = prop.multiProcessorCount;
  double* d = nullptr;
  CU_TRY(cudaMalloc((void**)&d, 64));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  float ms = 0.f;
  // DMMA: 2 CTAs/SM x 8 warps, 16 chains per warp
  {
    const int iters = 4096, grid = sms * 2;
    k_peak_dmma<<<grid, 256>>>(d, 64);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
      CU_TRY(cudaEventRecord(e0));
      k_peak_dmma<<<grid, 256>>>(d, iters);
      CU_TRY(cudaEventRecord(e1));
      CU_TRY(cudaEventSynchronize(e1));
      CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
      const double fl = (double)grid * 8 * iters * 16 * 512.0;
      best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    out_host[0] = best;
  }
  {
    const int iters = 4096, grid = sms * 4;
    k_peak_dfma<<<grid, 256>>>(d, 64);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
      CU_TRY(cudaEventRecord(e0));
      k_peak_dfma<<<grid, 256>>>(d, iters);
      CU_TRY(cudaEventRecord(e1));
      CU_TRY(cudaEventSynchronize(e1));
      CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
      const double fl = (double)grid * 256 * iters * 16 * 2.0;
      best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    out_host[1] = best;
  }
  {
    if (bytes < (size_t)1 << 20) bytes = (size_t)1 << 20;
    double2* buf = nullptr;
    CU_TRY(cudaMalloc((void**)&buf, bytes));
    CU_TRY(cudaMemset(buf, 0, bytes));
    const size_t n2 = bytes / sizeof(double2);
    k_stream_read<<<sms * 4, 512>>>(buf, n2, d);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
      CU_TRY(cudaEventRecord(e0));
      k_stream_read<<<sms * 4, 512>>>(buf, n2, d);
      CU_TRY(cudaEventRecord(e1));
      CU_TRY(cudaEventSynchronize(e1));
      CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
      best = std::max(best, (double)bytes / (ms * 1e-3) / 1e9);
    }
    out_host[2] = best;
    cudaFree(buf);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  CU_TRY(cudaGetLastError());
  return OO_OK;
}

}  // extern "C"
