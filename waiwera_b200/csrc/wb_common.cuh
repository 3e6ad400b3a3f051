// wb_common.cuh -- context, error handling, pointer staging, timers.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/waiwera_b200.h"
#include "wb_eos.cuh"
#include "wb_state.cuh"

void wb_set_error(const char *fmt, ...);
// guards the per-context work-space registries (std::map keyed by context): contexts may live on different
// host threads (one per GPU); a context itself is not re-entrant
std::mutex &wb_registry_mutex();

// NCCL is bound at run time (dlopen of libnccl.so.2 on the first communicator call) so that the
// library shares whichever NCCL the host process already carries (PyTorch's bundled one, or the
// system's under an MPI/PETSc host) and single-GPU users need none at all.
struct WbNccl {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
};
const WbNccl *wb_nccl();  // nullptr (with wb_last_error set) if libnccl cannot be loaded

#define WB_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      wb_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return -1;                                                                           \
    }                                                                                      \
  } while (0)

// Synchronous copy that is complete when it returns.  cudaMemcpy from PAGEABLE host memory returns once the data sits
// in the driver's staging buffer -- the DMA to the device may still be in flight -- and every kernel of this library
// runs on a non-blocking stream, which is not ordered behind the legacy stream the copy uses: a kernel launched right
// after a set-up upload could read the destination before the data landed.
static inline cudaError_t wb_memcpy_sync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpy(dst, src, bytes, kind);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
  return e;
}

// cudaMemset of device memory is asynchronous with respect to the host and runs on the legacy stream: the same hazard
static inline cudaError_t wb_memset_sync(void *dst, int value, size_t bytes) {
  cudaError_t e = cudaMemset(dst, value, bytes);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
  return e;
}

#define WB_NCCL(call)                                                                      \
  do {                                                                                     \
    ncclResult_t r_ = (call);                                                              \
    if (r_ != ncclSuccess) {                                                               \
      wb_set_error("%s:%d NCCL error %s: %s", __FILE__, __LINE__, #call, wb_nccl()->GetErrorString(r_)); \
      return -2;                                                                           \
    }                                                                                      \
  } while (0)

#define WB_CHECK(cond, ...)    \
  do {                         \
    if (!(cond)) {             \
      wb_set_error(__VA_ARGS__); \
      return -3;               \
    }                          \
  } while (0)

#define WB_TRY(call)        \
  do {                      \
    int rc_ = (call);       \
    if (rc_ != 0) return rc_; \
  } while (0)

// number of SMs on a B200; grids of the persistent / grid-stride kernels are sized in multiples of it
#define WB_NUM_SMS 148

struct WbTimer {
  double ms = 0.0;
  int64_t count = 0;
};

struct WbHalo {
  int nneigh = 0;
  std::vector<int> rank;
  std::vector<int> send_ptr, recv_ptr;  // host CSR
  int32_t *d_send_idx = nullptr, *d_recv_idx = nullptr;
  int nsend = 0, nrecv = 0;
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr;  // (nsend|nrecv) * maxwidth
  int maxwidth = 0;
  bool recv_contiguous = false;  // ghost cell k of the receive list is local cell nowned + k
};

// ---- NVLink peer-to-peer exchange (one process per GPU, CUDA IPC mapped buffers) -------------------
// Every rank owns one comm region with the SAME layout; peers write into it directly over NVLink / NVSwitch
// and publish a sequence number afterwards, the owner spins on the sequence number (system-scope acquire).
// Used for the three per-iteration exchanges of GMRES -- ghost entries of x, the Gram-Schmidt
// coefficients, the norm -- so that none of them is a separate collective launch.
#define WB_P2P_MAX_RANKS 8
#define WB_P2P_MAXV 32
#define WB_P2P_SLOT_A 4096   // byte offsets inside the region
#define WB_P2P_SLOT_B (WB_P2P_SLOT_A + WB_P2P_MAX_RANKS * WB_P2P_MAXV * 8)
// the persistent (fused) GMRES kernel keeps its own, double-buffered slots and flag kinds 3..5 (wb_fused.cu)
// Its data travels in the "LL" format (as NCCL's low-latency protocol): every double is one 16-byte store of two words
// (low half | sequence << 32, high half | sequence << 32), so data and flag arrive together and no fence is needed.
#define WB_P2P_SLOT_FA (WB_P2P_SLOT_B + 1024)                                   // [2][ranks][MAXV] x 16 bytes
#define WB_P2P_SLOT_FB (WB_P2P_SLOT_FA + 2 * WB_P2P_MAX_RANKS * WB_P2P_MAXV * 16)  // [2][ranks] x 16 bytes
#define WB_P2P_GHOST (WB_P2P_SLOT_FB + 1024)
struct WbP2PDev {  // by-value kernel argument
  int on, rank, nranks;
  unsigned char *region[WB_P2P_MAX_RANKS];  // region[r] = rank r's region as mapped in this process (region[rank] local)
  int *err;                                 // device flag: set when a spin wait times out
};
struct WbHaloPush {  // what a kernel needs to push this rank's boundary entries to its neighbours
  const int32_t *idx;       // [nsend] local owned cell of every send entry
  const int32_t *dst_rank;  // [nsend] destination rank
  const int32_t *dst_off;   // [nsend] cell offset in the destination's ghost area
  const int32_t *nb_rank;   // [nneigh]
  int nsend, nneigh, width, seq;
};
struct WbP2P {
  bool on = false;
  void *local = nullptr;
  size_t bytes = 0;
  WbP2PDev dev = {};
  int seq_halo = 0, seq_a = 0, seq_b = 0;
  std::vector<int> recv_off;          // [nranks] cell offset in MY ghost area of the data rank r sends (or -1)
  std::vector<int> peer_off;          // [nneigh] cell offset in neighbour n's ghost area where my data goes
  int32_t *d_send_nb = nullptr;       // [nsend] neighbour slot of every send entry
  int32_t *d_nb_rank = nullptr;       // [nneigh]
  int32_t *d_nb_off = nullptr;        // [nneigh] = peer_off
  int32_t *d_nb_start = nullptr;      // [nneigh] = send_ptr
  int32_t *d_dst_rank = nullptr, *d_dst_off = nullptr;  // [nsend] destination of every send entry
  unsigned *d_counter = nullptr;
  // fused kernel: two LL ghost buffers (16 bytes per double) behind the plain ghost area
  size_t ll_off = 0, ll_stride = 0;   // byte offset of the first buffer in the region, bytes between the two
  std::vector<size_t> peer_ll_off, peer_ll_stride;  // [nranks] the same for every rank's region
  int *d_fseq = nullptr;              // [4] sequence numbers of the fused kernel's exchanges (halo, dots, norm), device resident
};
// flags live in the first 4 KB of the region: one 64-byte line per (kind, sender rank)
__host__ __device__ inline size_t wb_p2p_flag_off(int kind, int sender) { return (size_t)(kind * WB_P2P_MAX_RANKS + sender) * 64; }

struct wb_mat {
  wb_ctx *ctx = nullptr;
  int nb = 0, ncolb = 0, bs = 0, nnzb = 0;
  int32_t *d_rowptr = nullptr, *d_colidx = nullptr;
  double *d_val = nullptr;
  double *d_xloc = nullptr;  // (ncolb-nb)*bs: ghost entries of x (multi-GPU), filled by the halo exchange
  std::vector<int32_t> h_rowptr, h_colidx;
  bool owns = true;
  uint64_t version = 1;          // bumped whenever the values change (the sliced-ELL copy follows it)
  bool external_vals = false;    // the value array was handed out (wb_jacobian_pattern): assume it changes between calls
  struct WbSell *sell = nullptr; // sliced-ELL copy for the stand-alone SpMV (wb_linalg.cu), built on first use
  // TMA-staged SpMV: first block of every tile of WB_SPMV_TILE rows (+ end), stage capacity in blocks
  int32_t *d_tile_e0 = nullptr;
  int ntiles = 0, tile_cap = 0;
};
#define WB_SPMV_TILE 128
#define WB_PAD_BYTES 256  // slack after rowptr / colidx / val so that 16-byte-rounded bulk copies stay inside the allocation
int wb_mat_build_tiles(wb_mat *A);  // after h_rowptr is known
void wb_sell_free(wb_mat *A);       // drops the sliced-ELL copy (wb_linalg.cu)

struct wb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  wb_params prm;
  WbEosParams eos;
  int np = 0, nc = 0, nph = 0, dof = 0, nf = 0;  // nf: SoA state fields per cell

  // mesh
  int ncell = 0, ninterior = 0, nowned = 0, nface = 0;
  std::vector<int32_t> h_face_cells;
  std::vector<double> h_rock, h_face_geom, h_cell_geom;
  int32_t *d_face_cells = nullptr;
  double *d_face = nullptr;   // SoA [6][nface]: area, d1, d2, d12, gravn, k
  double *d_vol = nullptr;    // [ncell]
  double *d_rockp = nullptr;  // SoA [5][ncell]: porosity, density, specific heat, wet, dry conductivity
  // cell -> faces (owned cells), entries in ascending face order = the reference's scatter order
  int32_t *d_cf_ptr = nullptr, *d_cf_face = nullptr, *d_cf_other = nullptr, *d_cf_bpos = nullptr;
  int32_t *d_diagpos = nullptr;
  int maxdeg = 0, ncf = 0;
  std::vector<int32_t> h_cf_ptr, h_cf_face, h_cf_other;

  // Jacobian
  wb_mat J;
  std::vector<int32_t> h_color;
  int ncolor = 0;
  std::vector<int32_t> pc_blocks;  // optional sub-domain assignment for the Newton solve's block Jacobi

  // state
  int32_t *d_region = nullptr, *d_region_iter = nullptr, *d_region_step = nullptr;  // [ncell]
  double *d_T_iter = nullptr, *d_T_step = nullptr;                                   // [ncell]
  double *d_sat_step = nullptr;  // [nph][ncell] saturations at the last time step (region 3 keeps them)
  double *d_state = nullptr;     // [(np+1) variants][nf][ncell]
  double *d_Lvar = nullptr;      // [(np+1)][np][nowned]
  double *d_dx = nullptr;        // [np][ninterior] FD steps
  double *d_yloc = nullptr;      // [ninterior*np] primaries incl. partition ghosts
  double *d_balances = nullptr;  // [nowned*np] lhs of the last unperturbed evaluation
  int eval_variant = 0;          // state slot of the last wb_pre_eval (0 unperturbed, 1 perturbed scratch)
  // time-stepping method whose residual wb_residual_be / wb_jacobian_be / wb_newton_solve_be evaluate
  int method = 0;
  double dt_last = 0.0;
  double *d_lhs_last2 = nullptr;  // BDF2: lhs two steps back
  // fixed-rate sources / sinks, sorted by cell
  int nsrc = 0;
  int32_t *d_src_head = nullptr, *d_src_cell = nullptr, *d_src_comp = nullptr;
  double *d_src_rate = nullptr, *d_src_enth = nullptr;
  std::vector<int> h_src_order;  // sorted position -> input position
  int32_t *d_src_ctrl = nullptr;  // source controls, sorted like the sources (null: none)
  double *d_src_pi = nullptr, *d_src_pref = nullptr, *d_src_limit = nullptr;
  std::vector<int32_t> h_src_ctrl;  // host copies of the control arrays (sorted source order): setters edit and re-upload
  std::vector<double> h_src_pi, h_src_pref, h_src_limit;
  int32_t *d_src_ptab_n = nullptr;  // reference-pressure tables of sources on deliverability (WbSources::ptab_n / ptab)
  double *d_src_ptab = nullptr;
  int32_t *d_src_sep_n = nullptr;  // separators: stages per source, reference enthalpies, separated-flow limits
  double *d_src_sep_h = nullptr, *d_src_limit_w = nullptr, *d_src_limit_s = nullptr;
  // passive tracers: auxiliary linear problem (wb_tracer.cu)
  int nt = 0;
  int trc_phase[WB_MAX_TRACERS] = {0, 0, 0};  // 1-based phase index
  double trc_diffusion[WB_MAX_TRACERS] = {0, 0, 0}, trc_decay[WB_MAX_TRACERS] = {0, 0, 0},
         trc_activation[WB_MAX_TRACERS] = {0, 0, 0};
  double *d_trc_inj = nullptr;  // [nsrc*nt] injection rates in sorted source order
  wb_mat *A_aux = nullptr;      // BAIJ bs = nt on the Jacobian's block pattern
  double *d_trc_b = nullptr, *d_trc_x = nullptr, *d_trc_al = nullptr;  // [nowned*nt] work vectors
  wb_pc *trc_pc = nullptr;      // preconditioner of A_aux (symbolic part kept between steps)
  int trc_pc_type = -1, trc_pc_nblocks = 0;
  int *d_flags = nullptr;        // [8] device flags (error, changed_y, changed_search, ...)
  int *h_flags = nullptr;        // pinned mirror

  // scratch pools
  std::vector<void *> stage;  // device staging buffers (freed at destroy)
  double *d_red = nullptr;    // reduction scratch
  double *h_red = nullptr;    // pinned
  size_t red_cap = 0;

  // comm
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int64_t first_cell = 0, ncell_global = 0;
  WbHalo halo;
  WbP2P p2p;

  // instrumentation
  std::map<std::string, WbTimer> timers;
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

// ---- pointer staging: accept host or device arrays at the ABI ----------------
bool wb_is_device_ptr(const void *p);

// RAII-less helpers (ctx owns the temporaries until wb_stage_release)
struct WbStage {
  wb_ctx *ctx;
  std::vector<void *> tmp;
  struct Out { void *host; void *dev; size_t bytes; };
  std::vector<Out> outs;
  explicit WbStage(wb_ctx *c) : ctx(c) {}
  ~WbStage();
  // device view of an input array (copied if it lives on the host); nullptr stays nullptr
  template <class T> const T *in(const T *p, size_t n, int *rc) { return (const T *)in_(p, n * sizeof(T), rc); }
  // device view of an output (or in/out when `load`) array; copied back by finish()
  template <class T> T *out(T *p, size_t n, int *rc, bool load = false) { return (T *)out_(p, n * sizeof(T), rc, load); }
  int finish();  // copies outputs back and synchronises the stream
  const void *in_(const void *p, size_t bytes, int *rc);
  void *out_(void *p, size_t bytes, int *rc, bool load);
};

// ---- timers --------------------------------------------------------------------
struct WbScopedTimer {
  wb_ctx *ctx;
  const char *name;
  bool on;
  WbScopedTimer(wb_ctx *c, const char *n);
  ~WbScopedTimer();
};
extern bool g_wb_timers_enabled;

#define WB_LAUNCH(ctx) ((ctx)->launches++)

static inline int wb_grid(size_t n, int block) { return (int)((n + block - 1) / block); }

// internal cross-file API
int wb_halo_exchange(wb_ctx *ctx, double *vec, int width);  // vec[(ninterior)*width]: fills ghost entries
// SpMV halo: owned[idx]*scale -> neighbours; ghost[(cell-nowned)*width+k] <- neighbours
int wb_halo_exchange_ghost(wb_ctx *ctx, const double *owned, int width, const double *scale, double *ghost);
int wb_allreduce_sum(wb_ctx *ctx, double *dbuf, int n);     // in-stream, device buffer
// P2P halo push of owned[idx]*scale into the neighbours' ghost areas; returns the sequence number consumers wait for
int wb_p2p_halo_push(wb_ctx *ctx, const double *owned, int width, const double *scale, const int *done, int *seq);
// arguments for a push fused into another kernel; takes the next halo sequence number
WbHaloPush wb_p2p_halo_push_args(wb_ctx *ctx, int width);
int wb_allreduce_max_int(wb_ctx *ctx, int *dbuf, int n);
int wb_reduce_flags(wb_ctx *ctx, int nflags);               // device flags -> host (max over ranks)

int wb_spmv_launch(wb_mat *A, const double *d_x, double *d_y);  // device pointers, handles halo
// y = A (x*scale); optionally stores the scaled owned entries to xn; skipped when *done
int wb_spmv_fused(wb_mat *A, const double *d_x, const double *d_scale, double *d_xn, double *d_y, const int *done,
                  int prepushed_halo_seq = 0);

// device-pointer cores shared between the translation units (no staging, no flag check)
int wb_pre_eval_dev(wb_ctx *c, const double *d_y, bool unperturbed);
int wb_residual_be_dev(wb_ctx *c, const double *d_y, const double *d_lhs_last, double dt, bool unperturbed,
                       double *d_lhs, double *d_rhs, double *d_r);
int wb_jacobian_be_dev(wb_ctx *c, const double *d_y, const double *d_lhs_last, double dt, double fd_err,
                       double fd_umin, bool base_valid);
int wb_fluid_transitions_dev(wb_ctx *c, const double *d_y_old, double *d_search, double *d_y);
int wb_max_scaled_core(wb_ctx *c, const double *d_v, const double *d_s, double tol, int n, double *maxval,
                       int64_t *maxloc);
int wb_pc_apply_dev(wb_pc *pc, const double *d_r, double *d_z, const int *done = nullptr);
int wb_ksp_solve_dev(wb_mat *A, wb_pc *pc, const wb_ksp_opts *o, const double *d_b, double *d_x, int *its,
                     int *reason, double *rnorm);
int wb_vec_dot_host(wb_ctx *c, const double *d_a, const double *d_b, size_t n, double *out);
int wb_vec_axpby_dev(wb_ctx *c, double *z, double a, const double *x, double b, const double *y, size_t n);
extern "C" int wb_pc_refactor(wb_pc *pc);
void wb_linalg_release(wb_ctx *c);
void wb_flow_release(wb_ctx *c);
void wb_newton_release(wb_ctx *c);
void wb_tracer_release(wb_ctx *c);
WbSources wb_sources_args(const wb_ctx *c);
