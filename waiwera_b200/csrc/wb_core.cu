// wb_core.cu -- context life cycle, error reporting, pointer staging, timers,
// NCCL communicator and halo exchange.
#include <algorithm>
#include <dlfcn.h>
#include <stdarg.h>

#include "wb_common.cuh"

static thread_local char g_err[1024] = "";
bool g_wb_timers_enabled = true;

std::mutex &wb_registry_mutex() {
  static std::mutex m;
  return m;
}

void wb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const WbNccl *wb_nccl() {
  static WbNccl tbl;
  static int state = 0;  // 0 untried, 1 ok, -1 failed
  if (state == 1) return &tbl;
  if (state == -1) return nullptr;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    wb_set_error("cannot load libnccl.so.2: %s", dlerror());
    state = -1;
    return nullptr;
  }
  bool ok = true;
#define BIND(field, name)                                   \
  do {                                                      \
    *(void **)(&tbl.field) = dlsym(h, name);                \
    if (!tbl.field) ok = false;                             \
  } while (0)
  BIND(GetUniqueId, "ncclGetUniqueId");
  BIND(CommInitRank, "ncclCommInitRank");
  BIND(CommDestroy, "ncclCommDestroy");
  BIND(GroupStart, "ncclGroupStart");
  BIND(GroupEnd, "ncclGroupEnd");
  BIND(Send, "ncclSend");
  BIND(Recv, "ncclRecv");
  BIND(AllReduce, "ncclAllReduce");
  BIND(AllGather, "ncclAllGather");
  BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
  if (!ok) {
    wb_set_error("libnccl.so.2 lacks a required symbol");
    state = -1;
    return nullptr;
  }
  state = 1;
  return &tbl;
}

extern "C" const char *wb_last_error(void) { return g_err; }
extern "C" int wb_version(void) { return WB_VERSION; }

bool wb_is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

const void *WbStage::in_(const void *p, size_t bytes, int *rc) {
  if (!p || bytes == 0) return p;
  if (wb_is_device_ptr(p)) return p;
  void *d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess) {
    wb_set_error("staging cudaMalloc(%zu) failed", bytes);
    *rc = -1;
    return nullptr;
  }
  tmp.push_back(d);
  if (cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
    wb_set_error("staging H2D copy failed");
    *rc = -1;
  }
  return d;
}

void *WbStage::out_(void *p, size_t bytes, int *rc, bool load) {
  if (!p || bytes == 0) return p;
  if (wb_is_device_ptr(p)) return p;
  void *d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess) {
    wb_set_error("staging cudaMalloc(%zu) failed", bytes);
    *rc = -1;
    return nullptr;
  }
  tmp.push_back(d);
  if (load && cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
    wb_set_error("staging H2D copy failed");
    *rc = -1;
  }
  outs.push_back({p, d, bytes});
  return d;
}

int WbStage::finish() {
  for (auto &o : outs) WB_CUDA(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx->stream));
  outs.clear();
  WB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

WbStage::~WbStage() {
  if (!tmp.empty()) {
    cudaStreamSynchronize(ctx->stream);
    for (void *d : tmp) cudaFree(d);
  }
}

WbScopedTimer::WbScopedTimer(wb_ctx *c, const char *n) : ctx(c), name(n), on(g_wb_timers_enabled) {
  if (on) cudaEventRecord(ctx->ev0, ctx->stream);
}
WbScopedTimer::~WbScopedTimer() {
  if (!on) return;
  cudaEventRecord(ctx->ev1, ctx->stream);
  cudaEventSynchronize(ctx->ev1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  WbTimer &t = ctx->timers[name];
  t.ms += ms;
  t.count++;
}

extern "C" int wb_timer_get(wb_ctx *ctx, const char *name, double *ms, int64_t *count) {
  auto it = ctx->timers.find(name);
  if (it == ctx->timers.end()) {
    if (ms) *ms = 0.0;
    if (count) *count = 0;
    return 0;
  }
  if (ms) *ms = it->second.ms;
  if (count) *count = it->second.count;
  return 0;
}
extern "C" int wb_timer_reset(wb_ctx *ctx) {
  ctx->timers.clear();
  return 0;
}
extern "C" int wb_timers_enable(int on) {
  g_wb_timers_enabled = on != 0;
  return 0;
}
extern "C" int64_t wb_launch_count(const wb_ctx *ctx) { return ctx->launches; }
extern "C" void *wb_stream(wb_ctx *ctx) { return (void *)ctx->stream; }

extern "C" int wb_create(const wb_params *prm, int device, wb_ctx **out) {
  WB_CHECK(prm && out, "wb_create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    // no CPU path exists: the engine is CUDA only
    wb_set_error("wb_create: no CUDA device available (%s); waiwera_b200 has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    cudaGetLastError();
    return -1;
  }
  WB_CHECK(device >= 0 && device < ndev, "wb_create: device %d out of range (%d devices)", device, ndev);
  WB_CUDA(cudaSetDevice(device));
  wb_ctx *c = new wb_ctx();
  c->device = device;
  c->prm = *prm;
  if (wb_eos_params_make(*prm, c->eos)) {
    delete c;
    wb_set_error("wb_create: unsupported eos id %d", prm->eos);
    return -3;
  }
  c->np = c->eos.np;
  c->nc = c->eos.nc;
  c->nph = c->eos.nphase;
  c->dof = 7 + c->nc - 1 + c->nph * (8 + c->nc - 1);  // src/fluid.F90:223-226
  c->nf = 4 + c->nph * (5 + (c->nc > 1 ? c->nc : 0));
  // from here on a failure releases what was created (wb_destroy copes with the partly built context)
  cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->ev0);
  if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->ev1);
  if (e2 == cudaSuccess) e2 = cudaMalloc(&c->d_flags, 8 * sizeof(int));
  if (e2 == cudaSuccess) e2 = wb_memset_sync(c->d_flags, 0, 8 * sizeof(int));
  if (e2 == cudaSuccess) e2 = cudaMallocHost(&c->h_flags, 8 * sizeof(int));
  c->red_cap = 64 * 1024;
  if (e2 == cudaSuccess) e2 = cudaMalloc(&c->d_red, c->red_cap * sizeof(double));
  if (e2 == cudaSuccess) e2 = cudaMallocHost(&c->h_red, 4096 * sizeof(double));
  if (e2 != cudaSuccess) {
    wb_set_error("wb_create: %s", cudaGetErrorString(e2));
    cudaGetLastError();
    c->J.ctx = c;
    wb_destroy(c);
    return -1;
  }
  c->J.ctx = c;
  *out = c;
  return 0;
}

static void free_mat(wb_mat &m) {
  if (m.owns) {
    cudaFree(m.d_rowptr);
    cudaFree(m.d_colidx);
    cudaFree(m.d_val);
  }
  cudaFree(m.d_xloc);
  cudaFree(m.d_tile_e0);
  wb_sell_free(&m);
  m.d_tile_e0 = nullptr;
  m.d_rowptr = m.d_colidx = nullptr;
  m.d_val = m.d_xloc = nullptr;
}

void wb_free_mesh(wb_ctx *c) {
  wb_tracer_release(c);  // A_aux borrows the Jacobian's pattern
  c->nt = 0;
  void *ptrs[] = {c->d_face_cells, c->d_face, c->d_vol, c->d_rockp, c->d_cf_ptr, c->d_cf_face, c->d_cf_other,
                  c->d_cf_bpos, c->d_diagpos, c->d_region, c->d_region_iter, c->d_region_step, c->d_T_iter,
                  c->d_T_step, c->d_sat_step, c->d_state, c->d_Lvar, c->d_dx, c->d_yloc, c->d_balances};
  for (void *p : ptrs) cudaFree(p);
  c->d_face_cells = c->d_cf_ptr = c->d_cf_face = c->d_cf_other = c->d_cf_bpos = c->d_diagpos = nullptr;
  c->d_region = c->d_region_iter = c->d_region_step = nullptr;
  c->d_face = c->d_vol = c->d_rockp = c->d_T_iter = c->d_T_step = c->d_sat_step = nullptr;
  c->d_state = c->d_Lvar = c->d_dx = c->d_yloc = c->d_balances = nullptr;
  free_mat(c->J);
}

extern "C" int wb_destroy(wb_ctx *c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  wb_newton_release(c);
  wb_linalg_release(c);
  wb_flow_release(c);
  wb_free_mesh(c);
  for (void *p : c->stage) cudaFree(p);
  cudaFree(c->d_lhs_last2);
  cudaFree(c->d_trc_inj);
  cudaFree(c->d_src_ctrl); cudaFree(c->d_src_pi); cudaFree(c->d_src_pref); cudaFree(c->d_src_limit);
  cudaFree(c->d_src_sep_n); cudaFree(c->d_src_sep_h); cudaFree(c->d_src_limit_w); cudaFree(c->d_src_limit_s);
  cudaFree(c->d_src_ptab_n); cudaFree(c->d_src_ptab);
  cudaFree(c->d_src_head); cudaFree(c->d_src_cell); cudaFree(c->d_src_comp); cudaFree(c->d_src_rate); cudaFree(c->d_src_enth);
  cudaFree(c->halo.d_send_idx);
  cudaFree(c->halo.d_recv_idx);
  cudaFree(c->halo.d_sendbuf);
  cudaFree(c->halo.d_recvbuf);
  if (c->p2p.local) {
    for (int r = 0; r < c->nranks; r++)
      if (r != c->rank && c->p2p.dev.region[r]) cudaIpcCloseMemHandle(c->p2p.dev.region[r]);
    cudaFree(c->p2p.local);
    cudaFree(c->p2p.d_send_nb); cudaFree(c->p2p.d_nb_rank); cudaFree(c->p2p.d_nb_off); cudaFree(c->p2p.d_nb_start);
    cudaFree(c->p2p.d_counter);
    cudaFree(c->p2p.d_dst_rank); cudaFree(c->p2p.d_dst_off);
    cudaFree(c->p2p.d_fseq);
  }
  if (c->comm && wb_nccl()) wb_nccl()->CommDestroy(c->comm);
  cudaFree(c->d_flags);
  cudaFreeHost(c->h_flags);
  cudaFree(c->d_red);
  cudaFreeHost(c->h_red);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

extern "C" int wb_num_primary(const wb_ctx *c) { return c->np; }
extern "C" int wb_fluid_dof(const wb_ctx *c) { return c->dof; }

// ---------------------------------------------------------------- comm

extern "C" int wb_comm_unique_id(void *id128) {
  ncclUniqueId id;
  if (!wb_nccl()) return -2;
  WB_NCCL(wb_nccl()->GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(id128, &id, 128);
  return 0;
}

extern "C" int wb_comm_init(wb_ctx *c, int rank, int nranks, const void *id128) {
  WB_CUDA(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  if (!wb_nccl()) return -2;
  WB_NCCL(wb_nccl()->CommInitRank(&c->comm, nranks, id, rank));
  c->rank = rank;
  c->nranks = nranks;
  return 0;
}

extern "C" int wb_set_global_offset(wb_ctx *c, int64_t first_cell, int64_t ncell_global) {
  c->first_cell = first_cell;
  c->ncell_global = ncell_global;
  return 0;
}

extern "C" int wb_set_halo(wb_ctx *c, int nneigh, const int32_t *neigh_rank, const int32_t *send_ptr,
                           const int32_t *send_idx, const int32_t *recv_ptr, const int32_t *recv_idx) {
  WB_CUDA(cudaSetDevice(c->device));
  WbHalo &h = c->halo;
  cudaFree(h.d_send_idx);
  cudaFree(h.d_recv_idx);
  cudaFree(h.d_sendbuf);
  cudaFree(h.d_recvbuf);
  h = WbHalo();
  h.nneigh = nneigh;
  if (nneigh == 0) return 0;
  h.rank.assign(neigh_rank, neigh_rank + nneigh);
  h.send_ptr.assign(send_ptr, send_ptr + nneigh + 1);
  h.recv_ptr.assign(recv_ptr, recv_ptr + nneigh + 1);
  h.nsend = send_ptr[nneigh];
  h.nrecv = recv_ptr[nneigh];
  h.maxwidth = WB_MAX_NP + 1;
  WB_CUDA(cudaMalloc(&h.d_send_idx, sizeof(int32_t) * (h.nsend + 1)));
  WB_CUDA(cudaMalloc(&h.d_recv_idx, sizeof(int32_t) * (h.nrecv + 1)));
  WB_CUDA(wb_memcpy_sync(h.d_send_idx, send_idx, sizeof(int32_t) * h.nsend, cudaMemcpyHostToDevice));
  WB_CUDA(wb_memcpy_sync(h.d_recv_idx, recv_idx, sizeof(int32_t) * h.nrecv, cudaMemcpyHostToDevice));
  // ghost cells are numbered owner by owner in the order they arrive: then the receive buffer is the ghost
  // part of the vector itself and the SpMV halo needs no unpack kernel
  h.recv_contiguous = true;
  for (int k = 0; k < h.nrecv; k++)
    if (recv_idx[k] != c->nowned + k) h.recv_contiguous = false;
  WB_CUDA(cudaMalloc(&h.d_sendbuf, sizeof(double) * (size_t)(h.nsend + 1) * h.maxwidth));
  WB_CUDA(cudaMalloc(&h.d_recvbuf, sizeof(double) * (size_t)(h.nrecv + 1) * h.maxwidth));
  return 0;
}

__global__ void k_halo_pack(const double *__restrict__ vec, const int32_t *__restrict__ idx, int n, int width,
                            double *__restrict__ buf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * width) {
    int c = i / width, k = i - c * width;
    buf[i] = vec[(size_t)idx[c] * width + k];
  }
}
__global__ void k_halo_unpack(double *__restrict__ vec, const int32_t *__restrict__ idx, int n, int width,
                              const double *__restrict__ buf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * width) {
    int c = i / width, k = i - c * width;
    vec[(size_t)idx[c] * width + k] = buf[i];
  }
}

// Ghost exchange of a cell vector with `width` doubles per cell (replaces
// DMGlobalToLocal / VecScatter, src/dm_utils.F90:480-498).  In-stream; no host sync.
int wb_halo_exchange(wb_ctx *c, double *vec, int width) {
  WbHalo &h = c->halo;
  if (c->nranks <= 1 || h.nneigh == 0) return 0;
  WB_CHECK(width <= h.maxwidth, "halo width %d too large", width);
  WB_CHECK(c->comm, "halo exchange without communicator");
  if (h.nsend > 0) {
    k_halo_pack<<<wb_grid((size_t)h.nsend * width, 256), 256, 0, c->stream>>>(vec, h.d_send_idx, h.nsend, width,
                                                                              h.d_sendbuf);
    WB_LAUNCH(c);
  }
  WB_NCCL(wb_nccl()->GroupStart());
  for (int n = 0; n < h.nneigh; n++) {
    int ns = h.send_ptr[n + 1] - h.send_ptr[n], nr = h.recv_ptr[n + 1] - h.recv_ptr[n];
    if (ns > 0)
      WB_NCCL(wb_nccl()->Send(h.d_sendbuf + (size_t)h.send_ptr[n] * width, (size_t)ns * width, ncclDouble, h.rank[n],
                       c->comm, c->stream));
    if (nr > 0)
      WB_NCCL(wb_nccl()->Recv(h.d_recvbuf + (size_t)h.recv_ptr[n] * width, (size_t)nr * width, ncclDouble, h.rank[n],
                       c->comm, c->stream));
  }
  WB_NCCL(wb_nccl()->GroupEnd());
  if (h.nrecv > 0) {
    k_halo_unpack<<<wb_grid((size_t)h.nrecv * width, 256), 256, 0, c->stream>>>(vec, h.d_recv_idx, h.nrecv, width,
                                                                                h.d_recvbuf);
    WB_LAUNCH(c);
  }
  return 0;
}

__global__ void k_halo_pack_scaled(const double *__restrict__ vec, const int32_t *__restrict__ idx, int n, int width,
                                   const double *scale, double *__restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * width) {
    const int c = i / width, k = i - c * width;
    buf[i] = vec[(size_t)idx[c] * width + k] * (scale ? *scale : 1.0);
  }
}
__global__ void k_halo_unpack_ghost(double *__restrict__ ghost, const int32_t *__restrict__ idx, int n, int width,
                                    int first_ghost, const double *__restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * width) {
    const int c = i / width, k = i - c * width;
    ghost[(size_t)(idx[c] - first_ghost) * width + k] = buf[i];
  }
}

// Ghost exchange for the SpMV (MatMult_MPIBAIJ's VecScatter): owned entries of `owned` (times the device
// scalar `scale`, if given) go to the neighbours; the entries of this rank's ghost cells arrive in
// ghost[(cell - nowned)*width + k].  The owned part is never copied.
int wb_halo_exchange_ghost(wb_ctx *c, const double *owned, int width, const double *scale, double *ghost) {
  WbHalo &h = c->halo;
  if (c->nranks <= 1 || h.nneigh == 0) return 0;
  WB_CHECK(width <= h.maxwidth, "halo width %d too large", width);
  WB_CHECK(c->comm, "halo exchange without communicator");
  if (h.nsend > 0) {
    k_halo_pack_scaled<<<wb_grid((size_t)h.nsend * width, 256), 256, 0, c->stream>>>(owned, h.d_send_idx, h.nsend,
                                                                                     width, scale, h.d_sendbuf);
    WB_LAUNCH(c);
  }
  double *rbuf = h.recv_contiguous ? ghost : h.d_recvbuf;
  WB_NCCL(wb_nccl()->GroupStart());
  for (int n = 0; n < h.nneigh; n++) {
    int ns = h.send_ptr[n + 1] - h.send_ptr[n], nr = h.recv_ptr[n + 1] - h.recv_ptr[n];
    if (ns > 0)
      WB_NCCL(wb_nccl()->Send(h.d_sendbuf + (size_t)h.send_ptr[n] * width, (size_t)ns * width, ncclDouble, h.rank[n],
                       c->comm, c->stream));
    if (nr > 0)
      WB_NCCL(wb_nccl()->Recv(rbuf + (size_t)h.recv_ptr[n] * width, (size_t)nr * width, ncclDouble, h.rank[n],
                       c->comm, c->stream));
  }
  WB_NCCL(wb_nccl()->GroupEnd());
  if (!h.recv_contiguous && h.nrecv > 0) {
    k_halo_unpack_ghost<<<wb_grid((size_t)h.nrecv * width, 256), 256, 0, c->stream>>>(ghost, h.d_recv_idx, h.nrecv,
                                                                                      width, c->nowned, h.d_recvbuf);
    WB_LAUNCH(c);
  }
  return 0;
}

// ================================================================ NVLink peer-to-peer exchange

struct WbP2PBlob {
  cudaIpcMemHandle_t handle;
  int recv_off[WB_P2P_MAX_RANKS];
  int ghost_cells;
  int pad;
  unsigned long long ll_off, ll_stride;
};

extern "C" int wb_comm_p2p_blob_size(void) { return (int)sizeof(WbP2PBlob); }

// after wb_comm_init + wb_set_halo: allocate this rank's comm region and describe it for the peers
extern "C" int wb_comm_p2p_export(wb_ctx *c, void *blob) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(c->nranks > 1 && c->nranks <= WB_P2P_MAX_RANKS, "wb_comm_p2p_export: needs 2..%d ranks", WB_P2P_MAX_RANKS);
  WbP2P &p = c->p2p;
  WbHalo &h = c->halo;
  WB_CHECK(h.recv_contiguous || h.nrecv == 0, "wb_comm_p2p_export: ghost cells must be numbered in receive order");
  const int nghost = c->ninterior - c->nowned;
  p.ll_off = WB_P2P_GHOST + ((((size_t)(nghost + 1) * h.maxwidth * sizeof(double)) + 255) & ~(size_t)255);
  p.ll_stride = (((size_t)(nghost + 1) * h.maxwidth * 16) + 255) & ~(size_t)255;
  p.bytes = p.ll_off + 2 * p.ll_stride;
  WB_CUDA(cudaMalloc(&p.local, p.bytes));
  WB_CUDA(wb_memset_sync(p.local, 0, p.bytes));
  WbP2PBlob b;
  memset(&b, 0, sizeof(b));
  WB_CUDA(cudaIpcGetMemHandle(&b.handle, p.local));
  p.recv_off.assign(c->nranks, -1);
  for (int n = 0; n < h.nneigh; n++) p.recv_off[h.rank[n]] = h.recv_ptr[n];
  for (int r = 0; r < WB_P2P_MAX_RANKS; r++) b.recv_off[r] = r < c->nranks ? p.recv_off[r] : -1;
  b.ghost_cells = nghost;
  b.ll_off = p.ll_off;
  b.ll_stride = p.ll_stride;
  memcpy(blob, &b, sizeof(b));
  return 0;
}

// blobs: nranks blobs in rank order (all-gathered by the host).  Maps every peer's region and turns the P2P path on.
extern "C" int wb_comm_p2p_open(wb_ctx *c, const void *blobs) {
  WB_CUDA(cudaSetDevice(c->device));
  WbP2P &p = c->p2p;
  WbHalo &h = c->halo;
  WB_CHECK(p.local, "wb_comm_p2p_open: call wb_comm_p2p_export first");
  const WbP2PBlob *B = (const WbP2PBlob *)blobs;
  memset(&p.dev, 0, sizeof(p.dev));
  for (int r = 0; r < c->nranks; r++) {
    if (r == c->rank) {
      p.dev.region[r] = (unsigned char *)p.local;
    } else {
      void *ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, B[r].handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        wb_set_error("wb_comm_p2p_open: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
        return -1;
      }
      p.dev.region[r] = (unsigned char *)ptr;
    }
  }
  p.peer_ll_off.assign(c->nranks, 0);
  p.peer_ll_stride.assign(c->nranks, 0);
  for (int r = 0; r < c->nranks; r++) {
    p.peer_ll_off[r] = (size_t)B[r].ll_off;
    p.peer_ll_stride[r] = (size_t)B[r].ll_stride;
  }
  p.dev.rank = c->rank;
  p.dev.nranks = c->nranks;
  p.dev.err = c->d_flags + 5;
  // where my data lands in each neighbour's ghost area
  p.peer_off.assign(h.nneigh, 0);
  std::vector<int32_t> send_nb(std::max(h.nsend, 1), 0), nb_rank(std::max(h.nneigh, 1), 0), nb_off(std::max(h.nneigh, 1), 0),
      nb_start(std::max(h.nneigh, 1), 0);
  for (int n = 0; n < h.nneigh; n++) {
    const int off = B[h.rank[n]].recv_off[c->rank];
    WB_CHECK(off >= 0, "wb_comm_p2p_open: rank %d does not expect data from rank %d", h.rank[n], c->rank);
    p.peer_off[n] = off;
    nb_rank[n] = h.rank[n];
    nb_off[n] = off;
    nb_start[n] = h.send_ptr[n];
    for (int k = h.send_ptr[n]; k < h.send_ptr[n + 1]; k++) send_nb[k] = n;
  }
  WB_CUDA(cudaMalloc(&p.d_send_nb, sizeof(int32_t) * send_nb.size()));
  WB_CUDA(cudaMalloc(&p.d_nb_rank, sizeof(int32_t) * nb_rank.size()));
  WB_CUDA(cudaMalloc(&p.d_nb_off, sizeof(int32_t) * nb_off.size()));
  WB_CUDA(cudaMalloc(&p.d_nb_start, sizeof(int32_t) * nb_start.size()));
  {
    std::vector<int32_t> dr(send_nb.size(), 0), doff(send_nb.size(), 0);
    for (int n = 0; n < h.nneigh; n++)
      for (int k = h.send_ptr[n]; k < h.send_ptr[n + 1]; k++) {
        dr[k] = h.rank[n];
        doff[k] = nb_off[n] + (k - h.send_ptr[n]);
      }
    WB_CUDA(cudaMalloc(&p.d_dst_rank, sizeof(int32_t) * dr.size()));
    WB_CUDA(cudaMalloc(&p.d_dst_off, sizeof(int32_t) * doff.size()));
    WB_CUDA(wb_memcpy_sync(p.d_dst_rank, dr.data(), sizeof(int32_t) * dr.size(), cudaMemcpyHostToDevice));
    WB_CUDA(wb_memcpy_sync(p.d_dst_off, doff.data(), sizeof(int32_t) * doff.size(), cudaMemcpyHostToDevice));
  }
  WB_CUDA(cudaMalloc(&p.d_counter, sizeof(unsigned)));
  WB_CUDA(wb_memset_sync(p.d_counter, 0, sizeof(unsigned)));
  WB_CUDA(cudaMalloc(&p.d_fseq, 4 * sizeof(int)));
  WB_CUDA(wb_memset_sync(p.d_fseq, 0, 4 * sizeof(int)));
  WB_CUDA(wb_memcpy_sync(p.d_send_nb, send_nb.data(), sizeof(int32_t) * send_nb.size(), cudaMemcpyHostToDevice));
  WB_CUDA(wb_memcpy_sync(p.d_nb_rank, nb_rank.data(), sizeof(int32_t) * nb_rank.size(), cudaMemcpyHostToDevice));
  WB_CUDA(wb_memcpy_sync(p.d_nb_off, nb_off.data(), sizeof(int32_t) * nb_off.size(), cudaMemcpyHostToDevice));
  WB_CUDA(wb_memcpy_sync(p.d_nb_start, nb_start.data(), sizeof(int32_t) * nb_start.size(), cudaMemcpyHostToDevice));
  p.dev.on = 1;
  p.on = true;
  return 0;
}

extern "C" int wb_comm_p2p_enabled(const wb_ctx *c) { return c->p2p.on ? 1 : 0; }
extern "C" int wb_comm_p2p_disable(wb_ctx *c) {
  c->p2p.on = false;
  c->p2p.dev.on = 0;
  return 0;
}

__device__ __forceinline__ void p2p_store_release_sys(int *p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ghost entries: every send entry goes straight into its neighbour's ghost area over NVLink; the last CTA to
// finish publishes the sequence number to every neighbour (after a system-scope fence)
__global__ void k_p2p_halo_push(const double *__restrict__ vec, const int32_t *__restrict__ idx,
                                const int32_t *__restrict__ send_nb, const int32_t *__restrict__ nb_rank,
                                const int32_t *__restrict__ nb_off, const int32_t *__restrict__ nb_start, int nsend,
                                int nneigh, int width, const double *scale, WbP2PDev P, int seq, unsigned *counter,
                                const int *done) {
  if (done && *done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nsend * width) {
    const int e = i / width, k = i - e * width;
    const int n = send_nb[e];
    double *ghost = reinterpret_cast<double *>(P.region[nb_rank[n]] + WB_P2P_GHOST);
    ghost[(size_t)(nb_off[n] + (e - nb_start[n])) * width + k] = vec[(size_t)idx[e] * width + k] * (scale ? *scale : 1.0);
  }
  __shared__ bool s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  if (s_last && threadIdx.x < nneigh) {
    __threadfence_system();
    int *flag = reinterpret_cast<int *>(P.region[nb_rank[threadIdx.x]] + wb_p2p_flag_off(0, P.rank));
    p2p_store_release_sys(flag, seq);
  }
}

int wb_p2p_halo_push(wb_ctx *c, const double *owned, int width, const double *scale, const int *done, int *seq) {
  WbP2P &p = c->p2p;
  WbHalo &h = c->halo;
  p.seq_halo++;
  *seq = p.seq_halo;
  if (h.nneigh == 0) return 0;
  WB_CHECK(h.nneigh <= 256, "too many halo neighbours");
  const int grid = std::max(1, wb_grid((size_t)h.nsend * width, 256));
  k_p2p_halo_push<<<grid, 256, 0, c->stream>>>(owned, h.d_send_idx, p.d_send_nb, p.d_nb_rank, p.d_nb_off, p.d_nb_start,
                                              h.nsend, h.nneigh, width, scale, p.dev, p.seq_halo, p.d_counter, done);
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

WbHaloPush wb_p2p_halo_push_args(wb_ctx *c, int width) {
  WbP2P &p = c->p2p;
  WbHalo &h = c->halo;
  WbHaloPush a = {h.d_send_idx, p.d_dst_rank, p.d_dst_off, p.d_nb_rank, h.nsend, h.nneigh, width, ++p.seq_halo};
  return a;
}

int wb_allreduce_sum(wb_ctx *c, double *dbuf, int n) {
  if (c->nranks <= 1) return 0;
  WB_NCCL(wb_nccl()->AllReduce(dbuf, dbuf, n, ncclDouble, ncclSum, c->comm, c->stream));
  return 0;
}

int wb_allreduce_max_int(wb_ctx *c, int *dbuf, int n) {
  if (c->nranks <= 1) return 0;
  WB_NCCL(wb_nccl()->AllReduce(dbuf, dbuf, n, ncclInt, ncclMax, c->comm, c->stream));
  return 0;
}

// device flags -> pinned host mirror, maximum over ranks (the reference's
// mpi_broadcast_error_flag / Allreduce(LOR), src/mpi_utils.F90:46), then cleared on device.
__global__ void k_clear_flags(int *f, int n) {
  if (threadIdx.x < n) f[threadIdx.x] = 0;
}
int wb_reduce_flags(wb_ctx *c, int nflags) {
  WB_TRY(wb_allreduce_max_int(c, c->d_flags, nflags));
  WB_CUDA(cudaMemcpyAsync(c->h_flags, c->d_flags, nflags * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  k_clear_flags<<<1, 32, 0, c->stream>>>(c->d_flags, nflags);
  WB_LAUNCH(c);
  WB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
