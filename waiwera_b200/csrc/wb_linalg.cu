// wb_linalg.cu -- the PETSc-side operators of the Newton step on the GPU:
// BAIJ SpMV (K5), point-block Jacobi and block-Jacobi/ILU(0) preconditioners (K6),
// GMRES / BiCGStab with fused multi-dot / multi-axpy (K7).
//
// These stand in for PETSc library code that is not in the reference tree
// (MatMult_SeqBAIJ_N, MatILUFactorNumeric_SeqBAIJ_N_NaturalOrdering, MatSolve,
// KSPSolve_GMRES, KSPSolve_BCGS; call sites src/timestepper.F90:1645-1836).
// Block storage is PETSc BAIJ: bs x bs blocks, column-major inside a block.
#include <algorithm>
#include <math.h>
#include <stdlib.h>

#include "wb_linalg.cuh"


// ================================================================ SpMV (K5)

// Eight lanes per block row: lane l takes blocks rowptr[i]+l, +8, ...  Consecutive rows are
// consecutive in `val`, so a warp streams one contiguous span of the value array (the 78 % of
// the algorithmic bytes) with every sector fully used; x is gathered through L2 (16 MB at
// 1 M cells, resident in the 126 MB L2); a 3-step shuffle folds the eight partial block
// products, in a fixed order.
//
// Fused Krylov form: with `scale` (device scalar) the operand is x*scale, rounded entry by entry
// exactly as if the normalised vector had been stored first (VecScale then MatMult), and with `xn`
// the row's own normalised entries are stored as a by-product -- this is how GMRES normalises the
// new basis vector without a separate pass.  Ghost columns (col >= nb, multi-GPU) are read from
// `xg`, which the halo exchange fills (MatMult_MPIBAIJ's off-diagonal part).
struct SpmvArgs {
  const int32_t *rowptr, *colidx;
  const double *val, *x, *xg;
  const double *scale;  // nullable
  double *xn;           // nullable: xn[row] = x[row]*scale
  double *y;
  int nb;
  const int *done;  // nullable: device-side "solver finished" flag
  // P2P halo: ghost entries are pushed by the neighbours; wait (lazily, on the first ghost column) for their
  // sequence numbers.  hseq == 0: no waiting (NCCL path or no ghosts)
  WbP2PDev P;
  int hseq, nwait;
  const int32_t *nb_rank;
};


// R block rows per 8-lane group, processed level by level (all row pointers, then all column indices and
// value blocks, then all x gathers): the three dependent global-memory round trips of a row are shared by R
// rows, which multiplies the bytes in flight per warp by R at the same occupancy.
template <int BS, int R, int MINB>
__global__ void __launch_bounds__(256, MINB) k_bsr_spmv(const SpmvArgs a) {
  if (a.done && *a.done) return;
  constexpr int B2 = BS * BS;
  const int lane = threadIdx.x & 7;
  const int group = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int gstride = gridDim.x * 32;
  const double s = a.scale ? *a.scale : 1.0;
  int e0[R], e1[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = group + r * gstride;
    e0[r] = 0; e1[r] = 0;
    if (row < a.nb) {
      e0[r] = a.rowptr[row] + lane;
      e1[r] = a.rowptr[row + 1];
    }
  }
  double acc[R][BS];
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int i = 0; i < BS; i++) acc[r][i] = 0.0;
  bool more = false, waited = (a.hseq == 0);
#pragma unroll
  for (int r = 0; r < R; r++) more = more || (e0[r] < e1[r]);
  while (more) {
    int col[R];
    double v[R][B2], xb[R][BS];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const bool on = e0[r] < e1[r];
      col[r] = on ? __ldcs(a.colidx + e0[r]) : -1;  // evict-first like the values: single-use stream
      if (BS == 2) {
        double2 p = make_double2(0.0, 0.0), q = make_double2(0.0, 0.0);
        if (on) {
          // two 128-bit evict-first loads per 2x2 block (one 256-bit load measured slower here: 87 vs 71 us)
          p = __ldcs(reinterpret_cast<const double2 *>(a.val + (size_t)e0[r] * 4));
          q = __ldcs(reinterpret_cast<const double2 *>(a.val + (size_t)e0[r] * 4) + 1);
        }
        v[r][0] = p.x; v[r][1] = p.y; v[r][2] = q.x; v[r][3] = q.y;
      } else {
#pragma unroll
        for (int q = 0; q < B2; q++) v[r][q] = on ? __ldcs(a.val + (size_t)e0[r] * B2 + q) : 0.0;
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      const bool own = col[r] < a.nb;
      const double sc = own ? s : 1.0;  // ghost entries arrive already scaled by their owner
#pragma unroll
      for (int j = 0; j < BS; j++) xb[r][j] = 0.0;
      if (col[r] >= 0) {
        if (!own && !waited) {  // rows without ghost columns never wait: interior work overlaps the exchange
          for (int n = 0; n < a.nwait; n++) p2p_wait(p2p_my_flag(a.P, 0, a.nb_rank[n]), a.hseq, a.P.err);
          waited = true;
        }
        const double *xp = own ? a.x + (size_t)col[r] * BS : a.xg + (size_t)(col[r] - a.nb) * BS;
        if (!own) {
#pragma unroll
          for (int j = 0; j < BS; j++) xb[r][j] = __ldcg(xp + j) * sc;  // written by a peer GPU: read at L2
        } else if (BS == 2) {
          const double2 x2 = *reinterpret_cast<const double2 *>(xp);
          xb[r][0] = x2.x * sc; xb[r][1] = x2.y * sc;
        } else {
#pragma unroll
          for (int j = 0; j < BS; j++) xb[r][j] = xp[j] * sc;
        }
      }
    }
    more = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
      for (int j = 0; j < BS; j++)
#pragma unroll
        for (int i = 0; i < BS; i++) acc[r][i] += v[r][j * BS + i] * xb[r][j];
      e0[r] += 8;
      more = more || (e0[r] < e1[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
#pragma unroll
    for (int off = 4; off > 0; off >>= 1)
#pragma unroll
      for (int i = 0; i < BS; i++) acc[r][i] += __shfl_down_sync(0xffffffffu, acc[r][i], off, 8);
    const int row = group + r * gstride;
    if (row < a.nb && lane == 0) {
      if (BS == 2) *reinterpret_cast<double2 *>(a.y + (size_t)row * 2) = make_double2(acc[r][0], acc[r][1]);
      else {
#pragma unroll
        for (int i = 0; i < BS; i++) a.y[(size_t)row * BS + i] = acc[r][i];
      }
      if (a.xn) {
#pragma unroll
        for (int i = 0; i < BS; i++) a.xn[(size_t)row * BS + i] = a.x[(size_t)row * BS + i] * s;
      }
    }
  }
}

// ================================================================ SpMV (K5), sliced-ELL
// One warp per slice of 32 rows, thread per row: all of a row's blocks (up to 8 per round) are requested before the
// first use -- column indices and value planes with streaming loads (read once), then the operand entries through
// L2 -- so a 7-point row costs one round trip to HBM and one to L2, with no row pointer to chase first.  Same fused
// Krylov form as k_bsr_spmv (operand x * scale, normalised copy of the row's own entries), same arithmetic per row
// except that the block products are summed in column order by one thread instead of by eight lanes and a shuffle tree.
struct SellArgs {
  const int4 *slice;
  const unsigned char *data;
  const double *x, *xg, *scale;
  double *xn, *y;
  int nslices, nb;
  const int *done;
};
template <int BS, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_sell_spmv(const SellArgs a) {
  if (a.done && *a.done) return;
  constexpr int B2 = BS * BS, PW = IluPlane<BS>::PW, NPL = IluPlane<BS>::NP, CH = BS >= 3 ? 4 : 8;
  const int lane = threadIdx.x & 31;
  const int s0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const double s = a.scale ? *a.scale : 1.0;
  for (int sl = s0; sl < a.nslices; sl += gridDim.x * (blockDim.x >> 5)) {
    const int4 S = a.slice[sl];
    const int n = S.y & 255, nk = S.y >> 8;
    const unsigned char *base = a.data + (size_t)S.z * 16;
    const int32_t *ip = reinterpret_cast<const int32_t *>(base) + lane;
    const double *vp = reinterpret_cast<const double *>(base + (size_t)nk * WB_SELL_SLICE * 4) + (size_t)lane * PW;
    double acc[BS];
#pragma unroll
    for (int i = 0; i < BS; i++) acc[i] = 0.0;
    for (int k0 = 0; k0 < nk; k0 += CH) {
      int col[CH];
      double v[CH][B2];
#pragma unroll
      for (int u = 0; u < CH; u++) {
        const int k = min(k0 + u, nk - 1);  // past the end: the last block again, its x is zero
        col[u] = __ldcs(ip + k * WB_SELL_SLICE);
        const double *bp = vp + (size_t)k * B2 * WB_SELL_SLICE;
#pragma unroll
        for (int q = 0; q < NPL; q++) {
          if (PW == 2) {
            const double2 t = __ldcs(reinterpret_cast<const double2 *>(bp + (size_t)q * WB_SELL_SLICE * 2));
            v[u][2 * q] = t.x;
            v[u][2 * q + 1] = t.y;
          } else {
            v[u][q] = __ldcs(bp + (size_t)q * WB_SELL_SLICE);
          }
        }
      }
#pragma unroll
      for (int h0 = 0; h0 < CH; h0 += CH / 2) {
        double x[CH / 2][BS];
#pragma unroll
        for (int uh = 0; uh < CH / 2; uh++) {
          const int u = h0 + uh;
          const bool on = k0 + u < nk, own = col[u] < a.nb;
          const double sc = own ? s : 1.0;  // ghost entries arrive already scaled by their owner
          const double *xp = own ? a.x + (size_t)col[u] * BS : a.xg + (size_t)(col[u] - a.nb) * BS;
          if (BS == 2 && own) {
            const double2 t = on ? *reinterpret_cast<const double2 *>(xp) : make_double2(0.0, 0.0);
            x[uh][0] = t.x * sc;
            x[uh][1] = t.y * sc;
          } else {
#pragma unroll
            for (int j = 0; j < BS; j++) x[uh][j] = on ? (own ? xp[j] : __ldcg(xp + j)) * sc : 0.0;
          }
        }
#pragma unroll
        for (int uh = 0; uh < CH / 2; uh++)
#pragma unroll
          for (int j = 0; j < BS; j++)
#pragma unroll
            for (int i = 0; i < BS; i++) acc[i] += v[h0 + uh][j * BS + i] * x[uh][j];
      }
    }
    if (lane < n) {
      const size_t row = (size_t)S.x + lane;
      if (BS == 2) *reinterpret_cast<double2 *>(a.y + row * 2) = make_double2(acc[0], acc[1]);
      else {
#pragma unroll
        for (int i = 0; i < BS; i++) a.y[row * BS + i] = acc[i];
      }
      if (a.xn) {
#pragma unroll
        for (int i = 0; i < BS; i++) a.xn[row * BS + i] = a.x[row * BS + i] * s;
      }
    }
  }
}

void wb_sell_free(wb_mat *A) {
  WbSell *S = A->sell;
  if (!S) return;
  cudaFree(S->d_slice); cudaFree(S->d_ssrc); cudaFree(S->d_slot0); cudaFree(S->d_data);
  delete S;
  A->sell = nullptr;
}

static int sell_mode() {
  static int mode = -1;  // WB_SPMV_SELL = 0: keep the BAIJ kernel
  if (mode < 0) {
    const char *e = getenv("WB_SPMV_SELL");
    mode = e ? atoi(e) : 1;
  }
  return mode;
}

// symbolic part (once per pattern): slices of 32 consecutive rows in the matrix's own ordering
static int sell_build(wb_mat *A) {
  const int nb = A->nb, b2 = A->bs * A->bs;
  WbSell *S = new WbSell();
  std::vector<int4> slices;
  std::vector<int32_t> ssrc, slot0;
  std::vector<unsigned char> data;
  for (int r0 = 0; r0 < nb; r0 += WB_SELL_SLICE) {
    const int n = std::min(WB_SELL_SLICE, nb - r0);
    int nk = 1;
    for (int l = 0; l < n; l++) nk = std::max(nk, A->h_rowptr[r0 + l + 1] - A->h_rowptr[r0 + l]);
    const size_t bytes = (size_t)nk * (WB_SELL_SLICE * 4 + (size_t)b2 * WB_SELL_SLICE * 8), base = data.size();
    if (base + bytes >= ((size_t)1 << 35) || nk > 255) {
      delete S;
      return 1;
    }
    slices.push_back(make_int4(r0, n | (nk << 8), (int)(base / 16), (int)bytes));
    data.resize(base + bytes, 0);
    int32_t *ip = reinterpret_cast<int32_t *>(data.data() + base);
    const size_t sbase = ssrc.size();
    slot0.push_back((int32_t)sbase);
    ssrc.resize(sbase + (size_t)nk * WB_SELL_SLICE);
    for (int k = 0; k < nk; k++)
      for (int l = 0; l < WB_SELL_SLICE; l++) {
        const int row = r0 + std::min(l, n - 1);
        int col = row < A->ncolb ? row : 0, src = -1;  // padding: a zero block times the row's own entry
        if (l < n) {
          const int e = A->h_rowptr[row] + k;
          if (e < A->h_rowptr[row + 1]) {
            col = A->h_colidx[e];
            src = e;
          }
        }
        ip[(size_t)k * WB_SELL_SLICE + l] = col;
        ssrc[sbase + (size_t)k * WB_SELL_SLICE + l] = src;
      }
  }
  S->nslices = (int)slices.size();
  WB_TRY(upload(&S->d_slice, slices));
  WB_TRY(upload(&S->d_ssrc, ssrc));
  WB_TRY(upload(&S->d_slot0, slot0));
  WB_CUDA(cudaMalloc(&S->d_data, data.size() + WB_PAD_BYTES));
  WB_CUDA(wb_memcpy_sync(S->d_data, data.data(), data.size(), cudaMemcpyHostToDevice));
  A->sell = S;
  return 0;
}

int wb_sell_spmv(wb_mat *A, const double *d_x, const double *xg, const double *d_scale, double *d_xn, double *d_y,
                 const int *done) {
  wb_ctx *c = A->ctx;
  if (!sell_mode() || A->nb < 1024 || A->h_rowptr.empty()) return 1;  // tiny systems: not worth a second copy
  if (A->ncolb > A->nb && !xg) return 1;
  if (!A->sell) {
    const int rc = sell_build(A);
    if (rc) return rc;
  }
  WbSell *S = A->sell;
  if (S->version != A->version || A->external_vals) {  // numeric part: the values as they are now
    const int grid = wb_grid((size_t)S->nslices * 32, 256);
    switch (A->bs) {
      case 1: k_sell_fill<1><<<grid, 256, 0, c->stream>>>(A->d_val, S->d_slice, S->nslices, S->d_ssrc, S->d_slot0, S->d_data); break;
      case 2: k_sell_fill<2><<<grid, 256, 0, c->stream>>>(A->d_val, S->d_slice, S->nslices, S->d_ssrc, S->d_slot0, S->d_data); break;
      default: k_sell_fill<3><<<grid, 256, 0, c->stream>>>(A->d_val, S->d_slice, S->nslices, S->d_ssrc, S->d_slot0, S->d_data); break;
    }
    WB_LAUNCH(c);
    S->version = A->version;
  }
  SellArgs a = {S->d_slice, S->d_data, d_x, xg, d_scale, d_xn, d_y, S->nslices, A->nb, done};
  static int cta_threads = 0;  // tuning knob WB_SELL_CTA = 128 | 256 (threads per CTA; 512 threads per SM either way)
  if (!cta_threads) {
    const char *e = getenv("WB_SELL_CTA");
    cta_threads = e ? atoi(e) : 128;
    if (cta_threads != 128 && cta_threads != 256) cta_threads = 128;
  }
  const int grid = wb_grid((size_t)S->nslices, cta_threads / 32);  // one slice per warp
#define SELL(BS)                                                            \
  do {                                                                      \
    if (cta_threads == 128) k_sell_spmv<BS, 128><<<grid, 128, 0, c->stream>>>(a); \
    else k_sell_spmv<BS, 256><<<grid, 256, 0, c->stream>>>(a);              \
  } while (0)
  switch (A->bs) {
    case 1: SELL(1); break;
    case 2: SELL(2); break;
    default: SELL(3); break;
  }
#undef SELL
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

static bool wb_spmv_tma_enabled(const wb_mat *A);
static int wb_spmv_tma_launch(wb_mat *A, const SpmvArgs &a);

// y = A (x*scale) with device pointers; x holds the nb*bs owned entries.  With ghost columns
// (multi-GPU) the ghost entries are gathered straight into the matrix's ghost buffer by the halo
// exchange (MatMult_MPIBAIJ's VecScatter); the owned part is never copied.
int wb_spmv_fused(wb_mat *A, const double *d_x, const double *d_scale, double *d_xn, double *d_y, const int *done,
                  int prepushed_halo_seq) {
  wb_ctx *c = A->ctx;
  const double *xg = nullptr;
  int hseq = 0;
  if (A->ncolb > A->nb && c->nranks > 1) {
    if (c->p2p.on && A == &c->J) {
      // neighbours write their entries straight into this GPU's ghost area over NVLink (already done by the
      // tail of the previous multi-axpy when prepushed_halo_seq is set)
      if (prepushed_halo_seq > 0) hseq = prepushed_halo_seq;
      else WB_TRY(wb_p2p_halo_push(c, d_x, A->bs, d_scale, done, &hseq));
      xg = reinterpret_cast<const double *>(c->p2p.dev.region[c->rank] + WB_P2P_GHOST);
    } else {
      WB_TRY(wb_halo_exchange_ghost(c, d_x, A->bs, d_scale, A->d_xloc));
      xg = A->d_xloc;
    }
  } else if (A->ncolb > A->nb) {
    xg = A->d_xloc;  // ghost columns without a communicator: zeros
  }
  SpmvArgs a = {A->d_rowptr, A->d_colidx, A->d_val, d_x, xg, d_scale, d_xn, d_y, A->nb, done,
                c->p2p.dev, hseq, hseq ? c->halo.nneigh : 0, c->p2p.d_nb_rank};
  if (wb_spmv_tma_enabled(A)) return wb_spmv_tma_launch(A, a);
  if (hseq == 0) {
    const int rc = wb_sell_spmv(A, d_x, xg, d_scale, d_xn, d_y, done);
    if (rc <= 0) return rc;  // done (0) or failed (< 0); 1: not available for this matrix
  }
  // tuning knob (WB_SPMV_ROWS = 1, 2, 4).  R = 1 needs 32 registers => 8 CTAs / SM (full occupancy) and is the
  // fastest: 68.5 us vs 75.9 (R = 2, 54 registers) vs 97 (R = 4) at 1 M cells on the same box
  static int rows_per_group = 0;
  if (!rows_per_group) {
    const char *e = getenv("WB_SPMV_ROWS");
    rows_per_group = e ? atoi(e) : 1;
    if (rows_per_group != 1 && rows_per_group != 2 && rows_per_group != 4 && rows_per_group != 11 && rows_per_group != 12)
      rows_per_group = 1;
  }
  const int R = rows_per_group;  // 11 / 12: R = 1 / 2 with the register count capped for 8 / 6 CTAs per SM
  const int grid = std::max(1, wb_grid((size_t)A->nb, 32 * (R > 10 ? R - 10 : R)));
#define SPMV(BS)                                                                       \
  do {                                                                                 \
    if (R == 1) k_bsr_spmv<BS, 1, 1><<<grid, 256, 0, c->stream>>>(a);                  \
    else if (R == 11) k_bsr_spmv<BS, 1, 8><<<grid, 256, 0, c->stream>>>(a);            \
    else if (R == 2) k_bsr_spmv<BS, 2, 1><<<grid, 256, 0, c->stream>>>(a);             \
    else if (R == 12) k_bsr_spmv<BS, 2, 6><<<grid, 256, 0, c->stream>>>(a);            \
    else k_bsr_spmv<BS, 4, 1><<<grid, 256, 0, c->stream>>>(a);                         \
  } while (0)
  switch (A->bs) {
    case 1: SPMV(1); break;
    case 2: SPMV(2); break;
    case 3: SPMV(3); break;
    default: WB_CHECK(false, "wb_mat_mult: block size %d not supported", A->bs);
  }
#undef SPMV
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

int wb_spmv_launch(wb_mat *A, const double *d_x, double *d_y) {
  return wb_spmv_fused(A, d_x, nullptr, nullptr, d_y, nullptr);
}

extern "C" int wb_mat_create(wb_ctx *c, int nb, int ncolb, int bs, int nnzb, const int32_t *rowptr,
                             const int32_t *colidx, const double *vals, wb_mat **out) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(bs >= 1 && bs <= 3, "wb_mat_create: block size %d not supported", bs);
  WB_CHECK(ncolb >= nb, "wb_mat_create: ncolb < nb");
  wb_mat *A = new wb_mat();
  A->ctx = c; A->nb = nb; A->ncolb = ncolb; A->bs = bs; A->nnzb = nnzb; A->owns = true;
  A->h_rowptr.resize(nb + 1);
  A->h_colidx.resize(nnzb);
  WB_CUDA(wb_memcpy_sync(A->h_rowptr.data(), rowptr, sizeof(int32_t) * (nb + 1), cudaMemcpyDefault));
  WB_CUDA(wb_memcpy_sync(A->h_colidx.data(), colidx, sizeof(int32_t) * nnzb, cudaMemcpyDefault));
  WB_CUDA(cudaMalloc(&A->d_rowptr, sizeof(int32_t) * (nb + 1) + WB_PAD_BYTES));
  WB_CUDA(cudaMalloc(&A->d_colidx, sizeof(int32_t) * std::max(nnzb, 1) + WB_PAD_BYTES));
  WB_CUDA(cudaMalloc(&A->d_val, sizeof(double) * std::max<size_t>((size_t)nnzb * bs * bs, 1) + WB_PAD_BYTES));
  WB_CUDA(cudaMalloc(&A->d_xloc, sizeof(double) * (size_t)(ncolb - nb + 1) * bs));
  WB_CUDA(wb_memset_sync(A->d_xloc, 0, sizeof(double) * (size_t)(ncolb - nb + 1) * bs));
  WB_CUDA(wb_memcpy_sync(A->d_rowptr, A->h_rowptr.data(), sizeof(int32_t) * (nb + 1), cudaMemcpyHostToDevice));
  WB_CUDA(wb_memcpy_sync(A->d_colidx, A->h_colidx.data(), sizeof(int32_t) * nnzb, cudaMemcpyHostToDevice));
  if (vals) WB_CUDA(wb_memcpy_sync(A->d_val, vals, sizeof(double) * (size_t)nnzb * bs * bs, cudaMemcpyDefault));
  else WB_CUDA(wb_memset_sync(A->d_val, 0, sizeof(double) * (size_t)nnzb * bs * bs));
  WB_TRY(wb_mat_build_tiles(A));
  *out = A;
  return 0;
}

extern "C" int wb_mat_set_values(wb_mat *A, const double *vals) {
  WB_CUDA(cudaSetDevice(A->ctx->device));
  A->version++;
  WB_CUDA(cudaMemcpyAsync(A->d_val, vals, sizeof(double) * (size_t)A->nnzb * A->bs * A->bs, cudaMemcpyDefault,
                          A->ctx->stream));
  WB_CUDA(cudaStreamSynchronize(A->ctx->stream));
  return 0;
}

extern "C" int wb_mat_get_values(wb_mat *A, double *vals) {
  WB_CUDA(cudaSetDevice(A->ctx->device));
  WB_CUDA(cudaMemcpyAsync(vals, A->d_val, sizeof(double) * (size_t)A->nnzb * A->bs * A->bs, cudaMemcpyDefault,
                          A->ctx->stream));
  WB_CUDA(cudaStreamSynchronize(A->ctx->stream));
  return 0;
}

extern "C" int wb_mat_destroy(wb_mat *A) {
  if (!A || !A->owns || A == &A->ctx->J) return 0;
  cudaSetDevice(A->ctx->device);
  cudaFree(A->d_rowptr);
  cudaFree(A->d_colidx);
  cudaFree(A->d_val);
  cudaFree(A->d_xloc);
  cudaFree(A->d_tile_e0);
  wb_sell_free(A);
  delete A;
  return 0;
}

extern "C" int wb_mat_mult(wb_mat *A, const double *x, double *y) {
  wb_ctx *c = A->ctx;
  WB_CUDA(cudaSetDevice(c->device));
  int rc = 0;
  WbStage st(c);
  const size_t n = (size_t)A->nb * A->bs;
  const double *dx = st.in(x, n, &rc);
  double *dy = st.out(y, n, &rc);
  if (rc) return rc;
  {
    WbScopedTimer tm(c, "mat_mult");
    WB_TRY(wb_spmv_launch(A, dx, dy));
  }
  return st.finish();
}

// ================================================================ small dense blocks

// Gauss-Jordan with partial pivoting on a bs x bs column-major block
template <int BS> __device__ __forceinline__ bool blk_invert(const double *a, double *inv) {
  double m[BS][2 * BS];
#pragma unroll
  for (int i = 0; i < BS; i++)
#pragma unroll
    for (int j = 0; j < BS; j++) {
      m[i][j] = a[j * BS + i];
      m[i][BS + j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll
  for (int cc = 0; cc < BS; cc++) {
    int piv = cc;
#pragma unroll
    for (int r = cc + 1; r < BS; r++)
      if (fabs(m[r][cc]) > fabs(m[piv][cc])) piv = r;
    // swap rows (static indexing to stay in registers)
#pragma unroll
    for (int r = cc + 1; r < BS; r++)
      if (r == piv) {
#pragma unroll
        for (int j = 0; j < 2 * BS; j++) {
          const double tmp = m[cc][j];
          m[cc][j] = m[r][j];
          m[r][j] = tmp;
        }
      }
    if (m[cc][cc] == 0.0) return false;
    const double d = 1.0 / m[cc][cc];
#pragma unroll
    for (int j = 0; j < 2 * BS; j++) m[cc][j] *= d;
#pragma unroll
    for (int r = 0; r < BS; r++)
      if (r != cc) {
        const double f = m[r][cc];
        if (f != 0.0) {
#pragma unroll
          for (int j = 0; j < 2 * BS; j++) m[r][j] -= f * m[cc][j];
        }
      }
  }
#pragma unroll
  for (int i = 0; i < BS; i++)
#pragma unroll
    for (int j = 0; j < BS; j++) inv[j * BS + i] = m[i][BS + j];
  return true;
}

template <int BS> __device__ __forceinline__ void blk_mul(const double *a, const double *b, double *cc) {
#pragma unroll
  for (int j = 0; j < BS; j++)
#pragma unroll
    for (int i = 0; i < BS; i++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < BS; k++) s += a[k * BS + i] * b[j * BS + k];
      cc[j * BS + i] = s;
    }
}
template <int BS> __device__ __forceinline__ void blk_mulsub(const double *a, const double *b, double *cc) {
#pragma unroll
  for (int j = 0; j < BS; j++)
#pragma unroll
    for (int i = 0; i < BS; i++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < BS; k++) s += a[k * BS + i] * b[j * BS + k];
      cc[j * BS + i] -= s;
    }
}

// ================================================================ preconditioners (K6)


template <int BS>
__global__ void k_pbjacobi_setup(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                 const double *__restrict__ val, double *__restrict__ dinv, int nb, int *flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  int d = -1;
  for (int e = rowptr[i]; e < rowptr[i + 1]; e++)
    if (colidx[e] == i) d = e;
  double a[BS * BS], inv[BS * BS];
#pragma unroll
  for (int q = 0; q < BS * BS; q++) a[q] = d >= 0 ? val[(size_t)d * BS * BS + q] : 0.0;
  if (!blk_invert<BS>(a, inv)) {
    atomicMax(&flags[3], 1);
    return;
  }
#pragma unroll
  for (int q = 0; q < BS * BS; q++) dinv[(size_t)i * BS * BS + q] = inv[q];
}

template <int BS>
__global__ void k_pbjacobi_apply(const double *__restrict__ dinv, const double *__restrict__ r,
                                 double *__restrict__ z, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  double rr[BS];
#pragma unroll
  for (int j = 0; j < BS; j++) rr[j] = r[(size_t)i * BS + j];
#pragma unroll
  for (int ii = 0; ii < BS; ii++) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < BS; j++) s += dinv[(size_t)i * BS * BS + j * BS + ii] * rr[j];
    z[(size_t)i * BS + ii] = s;
  }
}

__global__ void k_gather_vals(const double *__restrict__ src, const int32_t *__restrict__ map, int n, int bs2,
                              double *__restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * bs2) return;
  const int e = i / bs2, q = i - e * bs2;
  dst[i] = src[(size_t)map[e] * bs2 + q];
}

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// CTAs take tickets so that a running CTA only ever waits on rows owned by CTAs that started
// before it (rows are scheduled in dependency-level order): the spin waits cannot deadlock.
__device__ __forceinline__ int cta_ticket(int *ticket) {
  __shared__ int s_bid;
  if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1);
  __syncthreads();
  return s_bid;
}

// Block ILU(0), IKJ, natural ordering, inverted diagonal blocks kept in place
// (MatILUFactorNumeric_SeqBAIJ_N_NaturalOrdering).  One thread per row in level order; a row
// spins on the completion flag of each earlier row it eliminates with.  The arithmetic per
// row is that of the sequential loop, so the factors do not depend on the schedule.
template <int BS>
__global__ void __launch_bounds__(128) k_ilu0_factor(const int32_t *__restrict__ sched, int nsched,
                                                     const int32_t *__restrict__ rowptr,
                                                     const int32_t *__restrict__ colidx,
                                                     const int32_t *__restrict__ diag, double *val, int *flag,
                                                     int epoch, int *ticket, int *flags) {
  constexpr int B2 = BS * BS;
  const int t = cta_ticket(ticket) * blockDim.x + threadIdx.x;
  if (t >= nsched) return;
  const int i = sched[t];
  if (i < 0) return;
  const int r0 = rowptr[i], r1 = rowptr[i + 1], di = diag[i];
  for (int k = r0; k < di; k++) {
    const int kr = colidx[k];
    while (ld_acquire(&flag[kr]) != epoch) {
    }
    const int dk = diag[kr], k1 = rowptr[kr + 1];
    double aik[B2], dinv[B2], mult[B2];
#pragma unroll
    for (int q = 0; q < B2; q++) {
      aik[q] = val[(size_t)k * B2 + q];
      dinv[q] = __ldcg(&val[(size_t)dk * B2 + q]);
    }
    blk_mul<BS>(aik, dinv, mult);
#pragma unroll
    for (int q = 0; q < B2; q++) val[(size_t)k * B2 + q] = mult[q];
    for (int qq = dk + 1; qq < k1; qq++) {
      const int col = colidx[qq];
      int p = -1;
      for (int s = k + 1; s < r1; s++)
        if (colidx[s] == col) p = s;
      if (p >= 0) {
        double u[B2], tgt[B2];
#pragma unroll
        for (int q = 0; q < B2; q++) {
          u[q] = __ldcg(&val[(size_t)qq * B2 + q]);
          tgt[q] = val[(size_t)p * B2 + q];
        }
        blk_mulsub<BS>(mult, u, tgt);
#pragma unroll
        for (int q = 0; q < B2; q++) val[(size_t)p * B2 + q] = tgt[q];
      }
    }
  }
  double a[B2], inv[B2];
#pragma unroll
  for (int q = 0; q < B2; q++) a[q] = val[(size_t)di * B2 + q];
  if (!blk_invert<BS>(a, inv)) {
    atomicMax(&flags[3], 1);
#pragma unroll
    for (int q = 0; q < B2; q++) inv[q] = 0.0;
  }
#pragma unroll
  for (int q = 0; q < B2; q++) val[(size_t)di * B2 + q] = inv[q];
  __threadfence();
  st_release(&flag[i], epoch);
}

// forward (unit lower) and backward (inverted diagonal) block triangular solves
// (MatSolve_SeqBAIJ_N_NaturalOrdering), same scheduling scheme as the factorisation
template <int BS, bool FWD>
__global__ void __launch_bounds__(128) k_ilu0_solve(const int32_t *__restrict__ sched, int nsched,
                                                    const int32_t *__restrict__ rowptr,
                                                    const int32_t *__restrict__ colidx,
                                                    const int32_t *__restrict__ diag,
                                                    const double *__restrict__ val, const double *r, double *z,
                                                    int *flag, int epoch, int *ticket) {
  constexpr int B2 = BS * BS;
  const int t = cta_ticket(ticket) * blockDim.x + threadIdx.x;
  if (t >= nsched) return;
  const int i = sched[t];
  if (i < 0) return;
  const int di = diag[i];
  const int k0 = FWD ? rowptr[i] : di + 1, k1 = FWD ? di : rowptr[i + 1];
  double s[BS];
#pragma unroll
  for (int ii = 0; ii < BS; ii++) s[ii] = FWD ? r[(size_t)i * BS + ii] : z[(size_t)i * BS + ii];
  for (int k = k0; k < k1; k++) {
    const int col = colidx[k];
    double v[B2];
#pragma unroll
    for (int q = 0; q < B2; q++) v[q] = val[(size_t)k * B2 + q];
    while (ld_acquire(&flag[col]) != epoch) {
    }
    double xb[BS];
#pragma unroll
    for (int j = 0; j < BS; j++) xb[j] = __ldcg(&z[(size_t)col * BS + j]);
#pragma unroll
    for (int j = 0; j < BS; j++)
#pragma unroll
      for (int ii = 0; ii < BS; ii++) s[ii] -= v[j * BS + ii] * xb[j];
  }
  if (FWD) {
#pragma unroll
    for (int ii = 0; ii < BS; ii++) z[(size_t)i * BS + ii] = s[ii];
  } else {
    double tt[BS];
#pragma unroll
    for (int ii = 0; ii < BS; ii++) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < BS; j++) acc += val[(size_t)di * B2 + j * BS + ii] * s[j];
      tt[ii] = acc;
    }
#pragma unroll
    for (int ii = 0; ii < BS; ii++) z[(size_t)i * BS + ii] = tt[ii];
  }
  __threadfence();
  st_release(&flag[i], epoch);
}





struct IluSolveArgs {
  const int4 *blk, *lev;
  const int32_t *blk_rows;
  const double *stream, *r;
  double *z;
  int stage_words, nstage, desc_words;
  const int *done;
  long long *trace;  // debug: per CTA (start ns, end ns, SM id, 0) when non-null
};


#define ILU_ROWS_PER_THREAD 8
#define ILU_MAX_STAGES 8
// TMA variant: blockDim.x = consumer threads + one producer warp.  The producer warp's lane 0 refills ring slot
// st with level l+nstage as soon as the consumers have released it (empty[st] mbarrier), so issuing the bulk
// copies never sits on the consumers' level-to-level critical path; consumers synchronise among themselves on
// named barrier 1.
template <int BS, bool TMA>
__global__ void __launch_bounds__(288) k_ilu0_block_solve(const IluSolveArgs a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);       // [ILU_MAX_STAGES]
  uint64_t *empty = full + ILU_MAX_STAGES;                        // [ILU_MAX_STAGES]
  int4 *desc = reinterpret_cast<int4 *>(smem_raw + 128);  // (word offset, bytes, n, nk | bwd << 16) of every level
  double *ring = reinterpret_cast<double *>(smem_raw + 128) + a.desc_words;
  double *zs = ring + (size_t)a.nstage * a.stage_words;
  const int4 d = a.blk[blockIdx.x];
  const int row0 = d.x, nrows = d.y, lev0 = d.z, nl = d.w;
  const int tid = threadIdx.x;
  const int ncons = TMA ? (int)blockDim.x - 32 : (int)blockDim.x;
  if (a.trace && tid == 0) {
    long long t;
    unsigned sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    a.trace[4 * blockIdx.x] = t;
    a.trace[4 * blockIdx.x + 2] = sm;
  }
  if (TMA) {
    if (tid == 0) {
      for (int st = 0; st < a.nstage; st++) {
        mbar_init(&full[st], 1);
        mbar_init(&empty[st], 1);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int l = tid; l < nl; l += blockDim.x) desc[l] = a.lev[lev0 + l];
    __syncthreads();
    if (tid >= ncons) {
      // ---- producer warp
      if (tid == ncons) {
        int st = 0;
        uint32_t parity = 0;
        for (int l = 0; l < nl; l++) {
          if (l >= a.nstage) mbar_wait(&empty[st], parity ^ 1u);  // slot released by the consumers of level l-nstage
          const int4 L = desc[l];
          mbar_expect_tx(&full[st], (uint32_t)L.y);
          tma_load_1d(ring + (size_t)st * a.stage_words, a.stream + L.x, (uint32_t)L.y, &full[st]);
          if (++st == a.nstage) {
            st = 0;
            parity ^= 1u;
          }
        }
      }
      return;
    }
  }
  // ---- consumers.  Right-hand side of the sub-domain -> shared memory (overwritten in place by the sweeps); the
  // global row of each local row stays in registers for the final store when the sub-domain is small enough.
  // Slot `nrows` of the vector is the zero that padding blocks multiply.
  int grow[ILU_ROWS_PER_THREAD];
  const bool rows_in_regs = nrows <= ILU_ROWS_PER_THREAD * ncons;
  if (tid < BS) zs[nrows * BS + tid] = 0.0;
  for (int base = 0; base < nrows; base += ILU_ROWS_PER_THREAD * ncons) {
#pragma unroll
    for (int u = 0; u < ILU_ROWS_PER_THREAD; u++) {
      const int li = base + u * ncons + tid;
      grow[u] = li < nrows ? a.blk_rows[row0 + li] : -1;
    }
    double rv[ILU_ROWS_PER_THREAD][BS];
#pragma unroll
    for (int u = 0; u < ILU_ROWS_PER_THREAD; u++)
#pragma unroll
      for (int i = 0; i < BS; i++) rv[u][i] = grow[u] >= 0 ? a.r[(size_t)grow[u] * BS + i] : 0.0;
#pragma unroll
    for (int u = 0; u < ILU_ROWS_PER_THREAD; u++) {
      const int li = base + u * ncons + tid;
      if (li < nrows) {
#pragma unroll
        for (int i = 0; i < BS; i++) zs[li * BS + i] = rv[u][i];
      }
    }
  }
  bar_sync_named(1, ncons);
  if (TMA) {
    // Only ONE thread observes the arrival of a level's record: lane 0 of the last consumer warp (the warp with
    // the fewest rows -- none at all in most levels) waits for level l+1 while the others apply level l; the
    // consumer barrier that ends level l then orders the waiter's acquire before everybody's reads of level l+1,
    // which takes the ~100-cycle mbarrier round trip off the level-to-level critical path.
    int st = 0;
    uint32_t parity = 0;
    const bool waiter = tid == ncons - 32;
    if (nl > 0) mbar_wait(&full[0], 0);  // level 0: everybody
    for (int l = 0; l < nl; l++) {
      const int4 L = desc[l];
      ilu_level<BS>(ring + (size_t)st * a.stage_words, L.z, L.w & 0xffff, (L.w >> 16) != 0, zs, ncons, tid);
      int stn = st + 1;
      uint32_t parn = parity;
      if (stn == a.nstage) {
        stn = 0;
        parn ^= 1u;
      }
      if (waiter && l + 1 < nl) mbar_wait(&full[stn], parn);
      bar_sync_named(1, ncons);  // level l applied by all consumers: zs is consistent, the ring slot is free
      if (tid == 0) mbar_arrive(&empty[st]);
      st = stn;
      parity = parn;
    }
  } else {
    for (int l = 0; l < nl; l++) {
      const int4 L = a.lev[lev0 + l];
      ilu_level<BS>(a.stream + L.x, L.z, L.w & 0xffff, (L.w >> 16) != 0, zs, ncons, tid);
      bar_sync_named(1, ncons);
    }
  }
  for (int base = 0; base < nrows; base += ILU_ROWS_PER_THREAD * ncons) {
#pragma unroll
    for (int u = 0; u < ILU_ROWS_PER_THREAD; u++) {
      const int li = base + u * ncons + tid;
      if (li < nrows) {
        const int g = rows_in_regs ? grow[u] : a.blk_rows[row0 + li];
#pragma unroll
        for (int i = 0; i < BS; i++) a.z[(size_t)g * BS + i] = zs[li * BS + i];
      }
    }
  }
  if (a.trace && tid == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[4 * blockIdx.x + 1] = t;
  }
}

// numeric part: scatter the factor blocks into the value planes of the level stream
__global__ void k_ilu_repack(const double *__restrict__ fac, const int4 *__restrict__ map, int nmap, int b2,
                             double *__restrict__ stream) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nmap * b2) {
    const int e = i / b2, q = i - e * b2;
    const int4 m = map[e];  // (factor block, word offset of its copy in the level stream)
    stream[(size_t)m.y + q] = fac[(size_t)m.x * b2 + q];
  }
}

// host: level-ordered stream of every sub-domain (symbolic, once per pattern)
static int build_block_streams(wb_pc *pc, const std::vector<int32_t> &blk_of, const std::vector<int32_t> &rowptr,
                               const std::vector<int32_t> &colidx, const std::vector<int32_t> &diag) {
  const int nb = pc->nb, b2 = pc->bs * pc->bs;
  int nblk = 0;
  for (int i = 0; i < nb; i++) nblk = std::max(nblk, blk_of[i] + 1);
  std::vector<int32_t> bcount(nblk + 1, 0);
  for (int i = 0; i < nb; i++) bcount[blk_of[i] + 1]++;
  int maxrows = 0;
  for (int b = 0; b < nblk; b++) {
    maxrows = std::max(maxrows, bcount[b + 1]);
    bcount[b + 1] += bcount[b];
  }
  std::vector<int32_t> blk_rows(nb), local(nb);
  {
    std::vector<int32_t> fill(bcount.begin(), bcount.end() - 1);
    for (int i = 0; i < nb; i++) {
      local[i] = fill[blk_of[i]] - bcount[blk_of[i]];
      blk_rows[fill[blk_of[i]]++] = i;
    }
  }
  // levels per row within its sub-domain
  std::vector<int32_t> lf(nb, 0), lb(nb, 0);
  for (int i = 0; i < nb; i++) {
    int l = 0;
    for (int k = rowptr[i]; k < diag[i]; k++) l = std::max(l, lf[colidx[k]] + 1);
    lf[i] = l;
  }
  for (int i = nb - 1; i >= 0; i--) {
    int l = 0;
    for (int k = diag[i] + 1; k < rowptr[i + 1]; k++) l = std::max(l, lb[colidx[k]] + 1);
    lb[i] = l;
  }
  std::vector<int4> blk(nblk), lev;
  std::vector<double> stream;  // 8-byte words
  std::vector<int4> repack;
  std::vector<std::vector<int32_t>> rows_of_level;
  stream.reserve((size_t)pc->nnzb * b2 + (size_t)pc->nnzb / 2 + 4 * (size_t)nb);
  repack.reserve(pc->nnzb);
  int max_level_words = 0, max_level_rows = 0, max_levels = 0;
  for (int b = 0; b < nblk; b++) {
    const int r0 = bcount[b], nr = bcount[b + 1] - bcount[b];
    const int lev0 = (int)lev.size();
    for (int pass = 0; pass < 2; pass++) {
      const std::vector<int32_t> &lv = pass == 0 ? lf : lb;
      int nl = 0;
      for (int q = 0; q < nr; q++) nl = std::max(nl, lv[blk_rows[r0 + q]] + 1);
      rows_of_level.assign(nl, std::vector<int32_t>());
      if (pass == 0)
        for (int q = 0; q < nr; q++) rows_of_level[lv[blk_rows[r0 + q]]].push_back(blk_rows[r0 + q]);
      else
        for (int q = nr - 1; q >= 0; q--) rows_of_level[lv[blk_rows[r0 + q]]].push_back(blk_rows[r0 + q]);
      for (int l = 0; l < nl; l++) {
        // A level becomes one or more records: rows with no off-diagonal block in this sweep (e.g. all MINC matrix
        // cells in the backward sweep) go into records of their own without padding blocks, rows with 1..3 blocks
        // share the branch-free 3-block form, wider rows keep their exact width; a record is cut at ~16 KB so that
        // it streams through the shared-memory rings (rows of a level are independent: the cut is free).
        std::vector<int32_t> rows = rows_of_level[l];
        auto width = [&](int row) { return pass == 0 ? diag[row] - rowptr[row] : rowptr[row + 1] - diag[row] - 1; };
        auto cls = [&](int row) {
          const int w = width(row);
          return w == 0 ? 0 : (w <= 3 ? 3 : w);
        };
        std::stable_sort(rows.begin(), rows.end(), [&](int a_, int b_) { return cls(a_) < cls(b_); });
        size_t r_begin = 0;
        while (r_begin < rows.size()) {
          const int nk = cls(rows[r_begin]);
          const int stride = ilu_row_stride(pc->bs, nk, pass), ni4 = ilu_ni4(nk);
          const size_t max_rows = std::max<size_t>(32, 16384 / stride);
          size_t r_end = r_begin;
          while (r_end < rows.size() && cls(rows[r_end]) == nk && r_end - r_begin < max_rows) r_end++;
          const int n = (int)(r_end - r_begin);
          const size_t w0 = stream.size();
          const int words = n * stride / 8;  // stride is a multiple of 16 bytes
          WB_CHECK(w0 + words < ((size_t)1 << 31), "wb_pc_setup: factor stream too large for 32-bit word offsets");
          stream.resize(w0 + words, 0.0);
          const int zero_slot = nr * pc->bs * 8;  // byte offset of the zero entry behind the sub-domain vector
          for (int q = 0; q < n; q++) {
            const int row = rows[r_begin + q];
            int32_t *pidx =
                reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(&stream[w0]) + (size_t)q * stride);
            for (int e = 0; e < 4 * ni4; e++) pidx[e] = zero_slot;  // padding blocks multiply the zero slot
            pidx[0] = local[row] * pc->bs * 8;
            const size_t wblk = w0 + ((size_t)q * stride + (size_t)ni4 * 16) / 8;  // first block of the row (words)
            const int k0 = pass == 0 ? rowptr[row] : diag[row] + 1, k1 = pass == 0 ? diag[row] : rowptr[row + 1];
            for (int k = k0; k < k1; k++) {
              const int kk = k - k0;
              pidx[kk + 1] = local[colidx[k]] * pc->bs * 8;
              repack.push_back(make_int4(k, (int)(wblk + (size_t)kk * b2), 0, 0));
            }
            if (pass == 1) repack.push_back(make_int4(diag[row], (int)(wblk + (size_t)nk * b2), 0, 0));
          }
          int4 L;
          L.x = (int)w0; L.y = words * 8; L.z = n; L.w = nk | (pass << 16);
          lev.push_back(L);
          max_level_words = std::max(max_level_words, words);
          max_level_rows = std::max(max_level_rows, n);
          r_begin = r_end;
        }
      }
    }
    WB_CHECK((int)lev.size() - lev0 < (1 << 30), "wb_pc_setup: sub-domain with too many levels");
    blk[b].x = r0;
    blk[b].y = nr;
    blk[b].z = lev0;
    blk[b].w = (int)lev.size() - lev0;
    max_levels = std::max(max_levels, blk[b].w);
  }
  pc->nblk = nblk;
  pc->h_blk = blk;
  pc->h_lev = lev;
  pc->h_blk_rows = blk_rows;
  pc->max_block_rows = maxrows;
  pc->nrepack = (int)repack.size();
  pc->stream_words = stream.size();
  WB_TRY(upload(&pc->d_blk, blk));
  WB_TRY(upload(&pc->d_lev, lev));
  WB_TRY(upload(&pc->d_blk_rows, blk_rows));
  WB_TRY(upload(&pc->d_stream, stream));
  WB_TRY(upload(&pc->d_repack, repack));
  // ring geometry: 3 stages when that keeps >= 4 CTAs per SM (else 2); a level record must fit
  // the mbarrier transaction count (< 2^20 bytes); otherwise the stream is read from global memory
  pc->solve_threads = max_level_rows > 128 ? 256 : 128;
  pc->stage_words = max_level_words;
  pc->desc_words = 2 * max_levels;
  const size_t zs_bytes = (size_t)(maxrows + 1) * pc->bs * sizeof(double), stage_bytes = (size_t)max_level_words * 8;
  pc->nstage = 0;
  if (stage_bytes < (1u << 20)) {
    for (int ns = 3; ns >= 2; ns--) {
      const size_t need = 128 + (size_t)pc->desc_words * 8 + ns * stage_bytes + zs_bytes;
      if (need <= (ns == 2 ? 200u * 1024 : 55u * 1024)) {
        pc->nstage = ns;
        break;
      }
    }
  }
  return 0;
}

static void level_schedule(int nb, const std::vector<int32_t> &rowptr, const std::vector<int32_t> &colidx,
                           bool forward, std::vector<int32_t> &sched, int &nlev) {
  std::vector<int32_t> lev(nb, 0);
  nlev = 0;
  if (forward) {
    for (int i = 0; i < nb; i++) {
      int l = 0;
      for (int k = rowptr[i]; k < rowptr[i + 1]; k++)
        if (colidx[k] < i) l = std::max(l, lev[colidx[k]] + 1);
      lev[i] = l;
      nlev = std::max(nlev, l + 1);
    }
  } else {
    for (int i = nb - 1; i >= 0; i--) {
      int l = 0;
      for (int k = rowptr[i]; k < rowptr[i + 1]; k++)
        if (colidx[k] > i) l = std::max(l, lev[colidx[k]] + 1);
      lev[i] = l;
      nlev = std::max(nlev, l + 1);
    }
  }
  std::vector<int32_t> cnt(nlev + 1, 0);
  for (int i = 0; i < nb; i++) cnt[lev[i] + 1]++;
  // warp-aligned start of each level so no warp mixes dependent rows
  std::vector<int64_t> start(nlev + 1, 0);
  for (int l = 0; l < nlev; l++) start[l + 1] = start[l] + ((cnt[l + 1] + 31) / 32) * 32;
  sched.assign((size_t)start[nlev], -1);
  std::vector<int64_t> fill(start.begin(), start.end() - 1);
  if (forward) {
    for (int i = 0; i < nb; i++) sched[fill[lev[i]]++] = i;
  } else {
    for (int i = nb - 1; i >= 0; i--) sched[fill[lev[i]]++] = i;
  }
}


// additive Schwarz: rows of the residual -> the extended numbering; owned rows of the sub-domain solutions -> z
__global__ void k_asm_gather(const double *__restrict__ r, const int32_t *__restrict__ ext_row, int ne, int bs,
                             double *__restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)ne * bs) return;
  const int e = (int)(i / bs), k = (int)(i - (size_t)e * bs);
  out[i] = r[(size_t)ext_row[e] * bs + k];
}
__global__ void k_asm_scatter(const double *__restrict__ ze, const int32_t *__restrict__ own, int nb, int bs,
                              double *__restrict__ z) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nb * bs) return;
  const int row = (int)(i / bs), k = (int)(i - (size_t)row * bs);
  z[i] = ze[(size_t)own[row] * bs + k];
}

static bool g_skip_fused_plan = false;  // set while the inner PC of an additive Schwarz PC is built

static int pc_numeric(wb_pc *pc) {
  wb_mat *A = pc->A;
  wb_ctx *c = A->ctx;
  const int nb = pc->nb;
  if (pc->type == WB_PC_ASM_ILU0) {
    wb_mat *E = pc->asm_mat;
    const int bs2 = pc->bs * pc->bs;
    k_gather_vals<<<wb_grid((size_t)E->nnzb * bs2, 256), 256, 0, c->stream>>>(A->d_val, pc->d_asm_src, E->nnzb, bs2,
                                                                             E->d_val);
    WB_LAUNCH(c);
    E->version++;
    return pc_numeric(pc->asm_inner);
  }
  if (pc->type == WB_PC_PBJACOBI) {
    const int grid = wb_grid(nb, 128);
    switch (pc->bs) {
      case 1: k_pbjacobi_setup<1><<<grid, 128, 0, c->stream>>>(A->d_rowptr, A->d_colidx, A->d_val, pc->d_dinv, nb, c->d_flags); break;
      case 2: k_pbjacobi_setup<2><<<grid, 128, 0, c->stream>>>(A->d_rowptr, A->d_colidx, A->d_val, pc->d_dinv, nb, c->d_flags); break;
      default: k_pbjacobi_setup<3><<<grid, 128, 0, c->stream>>>(A->d_rowptr, A->d_colidx, A->d_val, pc->d_dinv, nb, c->d_flags); break;
    }
    WB_LAUNCH(c);
  } else if (pc->type == WB_PC_BJACOBI_ILU0) {
    const int bs2 = pc->bs * pc->bs;
    k_gather_vals<<<wb_grid((size_t)pc->nnzb * bs2, 256), 256, 0, c->stream>>>(A->d_val, pc->d_src, pc->nnzb, bs2,
                                                                              pc->d_val);
    WB_LAUNCH(c);
    pc->epoch++;
    WB_CUDA(cudaMemsetAsync(pc->d_ticket, 0, sizeof(int), c->stream));
    const int grid = wb_grid(pc->nsched_f, 128);
    switch (pc->bs) {
      case 1: k_ilu0_factor<1><<<grid, 128, 0, c->stream>>>(pc->d_sched_f, pc->nsched_f, pc->d_rowptr, pc->d_colidx, pc->d_diag, pc->d_val, pc->d_flag, pc->epoch, pc->d_ticket, c->d_flags); break;
      case 2: k_ilu0_factor<2><<<grid, 128, 0, c->stream>>>(pc->d_sched_f, pc->nsched_f, pc->d_rowptr, pc->d_colidx, pc->d_diag, pc->d_val, pc->d_flag, pc->epoch, pc->d_ticket, c->d_flags); break;
      default: k_ilu0_factor<3><<<grid, 128, 0, c->stream>>>(pc->d_sched_f, pc->nsched_f, pc->d_rowptr, pc->d_colidx, pc->d_diag, pc->d_val, pc->d_flag, pc->epoch, pc->d_ticket, c->d_flags); break;
    }
    WB_LAUNCH(c);
    if (pc->blocked) {
      k_ilu_repack<<<wb_grid((size_t)pc->nrepack * bs2, 256), 256, 0, c->stream>>>(pc->d_val, pc->d_repack, pc->nrepack,
                                                                                  bs2, pc->d_stream);
      WB_LAUNCH(c);
    }
  }
  WB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int wb_pc_destroy(wb_pc *pc) {
  if (!pc) return 0;
  cudaSetDevice(pc->A->ctx->device);
  cudaStreamSynchronize(pc->A->ctx->stream);
  void *ptrs[] = {pc->d_dinv, pc->d_rowptr, pc->d_colidx, pc->d_diag, pc->d_src, pc->d_sched_f, pc->d_sched_b,
                  pc->d_val, pc->d_flag, pc->d_ticket, pc->d_blk, pc->d_lev, pc->d_blk_rows, pc->d_stream,
                  pc->d_repack, pc->d_asm_src, pc->d_asm_row, pc->d_asm_own, pc->d_asm_r, pc->d_asm_z};
  for (void *p : ptrs) cudaFree(p);
  if (pc->asm_inner) wb_pc_destroy(pc->asm_inner);
  if (pc->asm_mat) wb_mat_destroy(pc->asm_mat);
  wb_fused_free(pc);
  delete pc;
  return 0;
}

// PCSetUp: symbolic part on the host (pattern restriction + level schedule), numeric part on the GPU
extern "C" int wb_pc_setup(wb_mat *A, int type, int nblocks, const int32_t *block_of_row, wb_pc **out) {
  wb_ctx *c = A->ctx;
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(type >= WB_PC_NONE && type <= WB_PC_ASM_ILU0, "wb_pc_setup: unknown type %d", type);
  wb_pc *pc = new wb_pc();
  pc->A = A; pc->type = type; pc->nb = A->nb; pc->bs = A->bs; pc->nblocks = std::max(nblocks, 1);
  const int nb = A->nb, bs2 = A->bs * A->bs;
  if (type == WB_PC_PBJACOBI) {
    WB_CUDA(cudaMalloc(&pc->d_dinv, sizeof(double) * (size_t)nb * bs2));
  } else if (type == WB_PC_ASM_ILU0) {
    // PCSetUp_ASM with overlap 1: every sub-domain grows by the columns of its rows (MatIncreaseOverlap), rows in
    // ascending order (sorted index sets); the sub-matrices A[E_b, E_b] are laid side by side and factored as the
    // block-Jacobi ILU(0) of that block-diagonal matrix, so the sub-domain resident solve is reused unchanged.
    std::vector<int32_t> blk(nb, 0);
    if (block_of_row) blk.assign(block_of_row, block_of_row + nb);
    else if (pc->nblocks > 1)
      for (int i = 0; i < nb; i++) blk[i] = (int)(((int64_t)i * pc->nblocks) / nb);
    int nblk = 0;
    for (int i = 0; i < nb; i++) nblk = std::max(nblk, blk[i] + 1);
    std::vector<std::vector<int32_t>> ext(nblk);
    for (int i = 0; i < nb; i++) {
      std::vector<int32_t> &e = ext[blk[i]];
      for (int k = A->h_rowptr[i]; k < A->h_rowptr[i + 1]; k++)
        if (A->h_colidx[k] < nb) e.push_back(A->h_colidx[k]);  // ghost columns: rows of another rank, not fetched
      e.push_back(i);
    }
    std::vector<int32_t> ext_row, ext_blk, own(nb, -1), rowptr(1, 0), colidx, src, loc(nb, -1);
    for (int b = 0; b < nblk; b++) {
      std::vector<int32_t> &e = ext[b];
      std::sort(e.begin(), e.end());
      e.erase(std::unique(e.begin(), e.end()), e.end());
      const int e0 = (int)ext_row.size();
      for (size_t q = 0; q < e.size(); q++) loc[e[q]] = e0 + (int)q;
      for (size_t q = 0; q < e.size(); q++) {
        const int i = e[q];
        ext_row.push_back(i);
        ext_blk.push_back(b);
        if (blk[i] == b) own[i] = e0 + (int)q;
        for (int k = A->h_rowptr[i]; k < A->h_rowptr[i + 1]; k++) {
          const int col = A->h_colidx[k];
          if (col < nb && loc[col] >= 0) {
            colidx.push_back(loc[col]);
            src.push_back(k);
          }
        }
        rowptr.push_back((int32_t)colidx.size());
      }
      for (int32_t i : e) loc[i] = -1;
      std::vector<int32_t>().swap(e);
    }
    pc->asm_ne = (int)ext_row.size();
    int rc = wb_mat_create(c, pc->asm_ne, pc->asm_ne, A->bs, (int)colidx.size(), rowptr.data(), colidx.data(), nullptr,
                           &pc->asm_mat);
    if (rc == 0) rc = upload(&pc->d_asm_src, src);
    if (rc == 0) rc = upload(&pc->d_asm_row, ext_row);
    if (rc == 0) rc = upload(&pc->d_asm_own, own);
    if (rc == 0 && cudaMalloc(&pc->d_asm_r, sizeof(double) * (size_t)pc->asm_ne * A->bs) != cudaSuccess) rc = 1;
    if (rc == 0 && cudaMalloc(&pc->d_asm_z, sizeof(double) * (size_t)pc->asm_ne * A->bs) != cudaSuccess) rc = 1;
    if (rc == 0) {
      const int bs2v = A->bs * A->bs;
      k_gather_vals<<<wb_grid((size_t)colidx.size() * bs2v, 256), 256, 0, c->stream>>>(A->d_val, pc->d_asm_src,
                                                                                      (int)colidx.size(), bs2v,
                                                                                      pc->asm_mat->d_val);
      g_skip_fused_plan = true;
      rc = wb_pc_setup(pc->asm_mat, WB_PC_BJACOBI_ILU0, nblk, ext_blk.data(), &pc->asm_inner);
      g_skip_fused_plan = false;
    }
    if (rc) {
      wb_pc_destroy(pc);
      return rc ? rc : 1;
    }
    *out = pc;
    return 0;
  } else if (type == WB_PC_BJACOBI_ILU0) {
    std::vector<int32_t> blk(nb, 0);
    if (block_of_row) blk.assign(block_of_row, block_of_row + nb);
    else if (pc->nblocks > 1) {
      // PETSc's default split: contiguous, equal-sized ranges of rows
      for (int i = 0; i < nb; i++) blk[i] = (int)(((int64_t)i * pc->nblocks) / nb);
    }
    std::vector<int32_t> rowptr(nb + 1, 0), colidx, src, diag(nb, -1);
    colidx.reserve(A->nnzb);
    src.reserve(A->nnzb);
    for (int i = 0; i < nb; i++) {
      for (int k = A->h_rowptr[i]; k < A->h_rowptr[i + 1]; k++) {
        const int col = A->h_colidx[k];
        if (col < nb && blk[col] == blk[i]) {
          if (col == i) diag[i] = (int32_t)colidx.size();
          colidx.push_back(col);
          src.push_back(k);
        }
      }
      rowptr[i + 1] = (int32_t)colidx.size();
      WB_CHECK(diag[i] >= 0, "wb_pc_setup: row %d has no diagonal block", i);
    }
    pc->nnzb = (int)colidx.size();
    std::vector<int32_t> sf, sb;
    level_schedule(nb, rowptr, colidx, true, sf, pc->nlev_f);
    level_schedule(nb, rowptr, colidx, false, sb, pc->nlev_b);
    pc->nsched_f = (int)sf.size();
    pc->nsched_b = (int)sb.size();
    WB_TRY(upload(&pc->d_rowptr, rowptr));
    WB_TRY(upload(&pc->d_colidx, colidx));
    WB_TRY(upload(&pc->d_diag, diag));
    WB_TRY(upload(&pc->d_src, src));
    WB_TRY(upload(&pc->d_sched_f, sf));
    WB_TRY(upload(&pc->d_sched_b, sb));
    WB_CUDA(cudaMalloc(&pc->d_val, sizeof(double) * (size_t)pc->nnzb * bs2));
    WB_CUDA(cudaMalloc(&pc->d_flag, sizeof(int) * nb));
    WB_CUDA(wb_memset_sync(pc->d_flag, 0, sizeof(int) * nb));
    WB_CUDA(cudaMalloc(&pc->d_ticket, sizeof(int)));
    // sub-domains small enough for one CTA's shared memory use the resident solve
    int nblk_used = 0, maxrows = 0;
    {
      std::vector<int32_t> cnt;
      for (int i = 0; i < nb; i++) {
        if ((int)cnt.size() <= blk[i]) cnt.resize(blk[i] + 1, 0);
        maxrows = std::max(maxrows, ++cnt[blk[i]]);
      }
      nblk_used = (int)cnt.size();
    }
    if (nblk_used > 1 && (size_t)maxrows * A->bs * sizeof(double) <= 160 * 1024) {
      WB_TRY(build_block_streams(pc, blk, rowptr, colidx, diag));
      pc->blocked = true;
      if (!g_skip_fused_plan)
        WB_TRY(wb_fused_build(pc, blk));  // leaves pc->fused null when the sub-domains do not fit the resident kernel
    }
  }
  int rc;
  {
    WbScopedTimer tm(c, "pc_setup");
    rc = pc_numeric(pc);
  }
  if (rc == 0) rc = wb_reduce_flags(c, 4);
  if (rc == 0 && c->h_flags[3]) {
    wb_set_error("wb_pc_setup: singular diagonal block");
    rc = 2;
  }
  if (rc) {
    wb_pc_destroy(pc);
    return rc;
  }
  *out = pc;
  return 0;
}

// PCSetUp again after the matrix values changed (same pattern): numeric factorisation only
extern "C" int wb_pc_refactor(wb_pc *pc) {
  wb_ctx *c = pc->A->ctx;
  WB_CUDA(cudaSetDevice(c->device));
  {
    WbScopedTimer tm(c, "pc_setup");
    WB_TRY(pc_numeric(pc));
  }
  WB_TRY(wb_reduce_flags(c, 4));
  if (c->h_flags[3]) {
    wb_set_error("wb_pc_refactor: singular diagonal block");
    return 2;
  }
  return 0;
}

int wb_pc_apply_dev(wb_pc *pc, const double *d_r, double *d_z, const int *done) {
  wb_ctx *c = pc->A->ctx;
  const int nb = pc->nb;
  if (pc->type == WB_PC_NONE) {
    if (d_r != d_z)
      WB_CUDA(cudaMemcpyAsync(d_z, d_r, sizeof(double) * (size_t)nb * pc->bs, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
  }
  if (pc->type == WB_PC_ASM_ILU0) {
    const size_t ne = (size_t)pc->asm_ne * pc->bs, n = (size_t)nb * pc->bs;
    k_asm_gather<<<wb_grid(ne, 256), 256, 0, c->stream>>>(d_r, pc->d_asm_row, pc->asm_ne, pc->bs, pc->d_asm_r);
    WB_LAUNCH(c);
    WB_TRY(wb_pc_apply_dev(pc->asm_inner, pc->d_asm_r, pc->d_asm_z, done));
    k_asm_scatter<<<wb_grid(n, 256), 256, 0, c->stream>>>(pc->d_asm_z, pc->d_asm_own, nb, pc->bs, d_z);
    WB_LAUNCH(c);
    WB_CUDA(cudaGetLastError());
    return 0;
  }
  if (pc->type == WB_PC_PBJACOBI) {
    const int grid = wb_grid(nb, 256);
    switch (pc->bs) {
      case 1: k_pbjacobi_apply<1><<<grid, 256, 0, c->stream>>>(pc->d_dinv, d_r, d_z, nb); break;
      case 2: k_pbjacobi_apply<2><<<grid, 256, 0, c->stream>>>(pc->d_dinv, d_r, d_z, nb); break;
      default: k_pbjacobi_apply<3><<<grid, 256, 0, c->stream>>>(pc->d_dinv, d_r, d_z, nb); break;
    }
    WB_LAUNCH(c);
    WB_CUDA(cudaGetLastError());
    return 0;
  }
  if (pc->blocked) {
    const int desc_words = pc->nstage > 0 ? pc->desc_words : 0;
    const size_t smem = 128 + (size_t)desc_words * 8 + (size_t)pc->nstage * pc->stage_words * 8 +
                        (size_t)(pc->max_block_rows + 1) * pc->bs * sizeof(double);
    IluSolveArgs a = {pc->d_blk, pc->d_lev, pc->d_blk_rows, pc->d_stream, d_r, d_z, pc->stage_words, pc->nstage,
                      desc_words, done, pc->d_trace};
#define BSOLVE(BS)                                                                                           \
  do {                                                                                                       \
    if (pc->nstage > 0) {                                                                                    \
      cudaFuncSetAttribute(k_ilu0_block_solve<BS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      k_ilu0_block_solve<BS, true><<<pc->nblk, pc->solve_threads + 32, smem, c->stream>>>(a);                \
    } else {                                                                                                 \
      cudaFuncSetAttribute(k_ilu0_block_solve<BS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      k_ilu0_block_solve<BS, false><<<pc->nblk, pc->solve_threads, smem, c->stream>>>(a);                    \
    }                                                                                                        \
  } while (0)
    switch (pc->bs) {
      case 1: BSOLVE(1); break;
      case 2: BSOLVE(2); break;
      default: BSOLVE(3); break;
    }
#undef BSOLVE
    WB_LAUNCH(c);
    WB_CUDA(cudaGetLastError());
    return 0;
  }
  // forward then backward sweep
  for (int pass = 0; pass < 2; pass++) {
    pc->epoch++;
    WB_CUDA(cudaMemsetAsync(pc->d_ticket, 0, sizeof(int), c->stream));
    const int ns = pass == 0 ? pc->nsched_f : pc->nsched_b;
    const int32_t *sched = pass == 0 ? pc->d_sched_f : pc->d_sched_b;
    const int grid = wb_grid(ns, 128);
#define SOLVE(BS, FWD)                                                                                          \
  k_ilu0_solve<BS, FWD><<<grid, 128, 0, c->stream>>>(sched, ns, pc->d_rowptr, pc->d_colidx, pc->d_diag, pc->d_val, \
                                                      d_r, d_z, pc->d_flag, pc->epoch, pc->d_ticket)
    if (pass == 0) {
      switch (pc->bs) {
        case 1: SOLVE(1, true); break;
        case 2: SOLVE(2, true); break;
        default: SOLVE(3, true); break;
      }
    } else {
      switch (pc->bs) {
        case 1: SOLVE(1, false); break;
        case 2: SOLVE(2, false); break;
        default: SOLVE(3, false); break;
      }
    }
#undef SOLVE
    WB_LAUNCH(c);
  }
  WB_CUDA(cudaGetLastError());
  return 0;
}

// debug / tuning aid (not part of the public header): one PC apply with a per-CTA timeline of the sub-domain
// solve; out[4*b] = start ns, end ns, SM id of sub-domain b.  Also lets the tuner override the ring depth.
extern "C" int wb_debug_pc_trace(wb_pc *pc, const double *d_r, double *d_z, long long *out, int nstage_override) {
  wb_ctx *c = pc->A->ctx;
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(pc->blocked, "wb_debug_pc_trace: not a sub-domain resident solve");
  if (nstage_override >= 0) pc->nstage = nstage_override;
  WB_CUDA(cudaMalloc(&pc->d_trace, sizeof(long long) * (4 * pc->nblk + 128)));
  WB_CUDA(wb_memset_sync(pc->d_trace, 0, sizeof(long long) * (4 * pc->nblk + 128)));
  int rc = wb_pc_apply_dev(pc, d_r, d_z, nullptr);
  cudaStreamSynchronize(c->stream);
  if (out) wb_memcpy_sync(out, pc->d_trace, sizeof(long long) * (4 * pc->nblk + 128), cudaMemcpyDeviceToHost);
  cudaFree(pc->d_trace);
  pc->d_trace = nullptr;
  return rc;
}

extern "C" int wb_pc_apply(wb_pc *pc, const double *r, double *z) {
  wb_ctx *c = pc->A->ctx;
  WB_CUDA(cudaSetDevice(c->device));
  int rc = 0;
  WbStage st(c);
  const size_t n = (size_t)pc->nb * pc->bs;
  const double *dr = st.in(r, n, &rc);
  double *dz = st.out(z, n, &rc);
  if (rc) return rc;
  {
    WbScopedTimer tm(c, "pc_apply");
    WB_TRY(wb_pc_apply_dev(pc, dr, dz, nullptr));
  }
  return st.finish();
}

// ================================================================ SpMV (K5), TMA-staged
// Persistent CTAs (2 per SM) walk tiles of WB_SPMV_TILE block rows.  A dedicated producer warp brings each
// tile's three contiguous spans -- row pointers, column indices, value blocks -- into a shared-memory ring
// with cp.async.bulk (TMA, mbarrier-completed, L2 evict-first), NSTAGE tiles ahead, so the only global
// round trip left on the consumers' path is the gather of x (L2 resident): the rowptr -> colidx -> x
// dependency chain of the plain kernel becomes shared -> shared -> L2.  Consumers use the same 8-lanes-per-row
// scheme and fixed-order shuffle fold as k_bsr_spmv (identical arithmetic), four rows per lane group in flight;
// each consumer warp releases a ring slot with one mbarrier arrive, no CTA-wide barrier.
#define SPMV_TMA_STAGES 3
struct SpmvTmaArgs {
  SpmvArgs a;
  const int32_t *tile_e0;  // [ntiles + 1]
  int ntiles, cap;         // cap: stage capacity in blocks
};

template <int BS, int SPMV_TMA_CONSUMERS>
__global__ void __launch_bounds__(SPMV_TMA_CONSUMERS + 32) k_bsr_spmv_tma(const SpmvTmaArgs t) {
  const SpmvArgs &a = t.a;
  if (a.done && *a.done) return;
  constexpr int B2 = BS * BS, TR = WB_SPMV_TILE, NG = SPMV_TMA_CONSUMERS / 8, R = TR / NG;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
  uint64_t *empty = full + SPMV_TMA_STAGES;
  // stage layout: [rowptr slice: (TR+4) ints][colidx: cap+8 ints][values: (cap+2)*B2 doubles], 16-byte aligned parts
  const int rp_bytes = (TR + 4) * 4, ci_bytes = ((t.cap + 8) * 4 + 15) & ~15, va_bytes = (t.cap + 4) * B2 * 8;
  const int stage_bytes = (rp_bytes + ci_bytes + va_bytes + 127) & ~127;
  unsigned char *stages = smem_raw + 128;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < SPMV_TMA_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], SPMV_TMA_CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid >= SPMV_TMA_CONSUMERS) {
    // ---- producer warp
    if (tid == SPMV_TMA_CONSUMERS) {
      int st = 0;
      uint32_t parity = 0;
      int k = 0;
      for (int tile = blockIdx.x; tile < t.ntiles; tile += gridDim.x, k++) {
        if (k >= SPMV_TMA_STAGES) mbar_wait(&empty[st], parity ^ 1u);
        const int r0 = tile * TR;
        const int nr = min(TR, a.nb - r0);
        const int e0 = t.tile_e0[tile], e1 = t.tile_e0[tile + 1];
        const int ec = e0 & ~3, ev = e0 & ~1;  // 16-byte aligned starts of the index / value spans
        const uint32_t b_rp = (uint32_t)(((nr + 1) * 4 + 15) & ~15);
        const uint32_t b_ci = (uint32_t)(((e1 - ec) * 4 + 15) & ~15);
        const uint32_t b_va = (uint32_t)((((size_t)(e1 - ev) * B2 * 8) + 15) & ~(size_t)15);
        unsigned char *sp = stages + (size_t)st * stage_bytes;
        mbar_expect_tx(&full[st], b_rp + b_ci + b_va);
        tma_load_1d(sp, a.rowptr + r0, b_rp, &full[st]);
        if (b_ci) tma_load_1d(sp + rp_bytes, a.colidx + ec, b_ci, &full[st]);
        if (b_va) tma_load_1d(sp + rp_bytes + ci_bytes, a.val + (size_t)ev * B2, b_va, &full[st]);
        if (++st == SPMV_TMA_STAGES) {
          st = 0;
          parity ^= 1u;
        }
      }
    }
    return;
  }
  // ---- consumers
  const int lane = tid & 7, grp = tid >> 3;  // NG lane groups, R rows each per tile
  const double s = a.scale ? *a.scale : 1.0;
  bool waited = (a.hseq == 0);
  int st = 0;
  uint32_t parity = 0;
  for (int tile = blockIdx.x; tile < t.ntiles; tile += gridDim.x) {
    const int r0 = tile * TR;
    const unsigned char *sp = stages + (size_t)st * stage_bytes;
    const int *rp = reinterpret_cast<const int *>(sp);
    const int *sci = reinterpret_cast<const int *>(sp + rp_bytes);
    const double *sva = reinterpret_cast<const double *>(sp + rp_bytes + ci_bytes);
    mbar_wait(&full[st], parity);
    const int e_base = rp[0];
    const int ec = e_base & ~3, ev = e_base & ~1;
    int e0[R], e1[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int rl = grp + r * NG;
      e0[r] = 0; e1[r] = 0;
      if (r0 + rl < a.nb) {
        e0[r] = rp[rl] + lane;
        e1[r] = rp[rl + 1];
      }
    }
    double acc[R][BS];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
      for (int i = 0; i < BS; i++) acc[r][i] = 0.0;
    bool more = false;
#pragma unroll
    for (int r = 0; r < R; r++) more = more || (e0[r] < e1[r]);
    while (more) {
      int col[R];
      double xb[R][BS];
#pragma unroll
      for (int r = 0; r < R; r++) col[r] = (e0[r] < e1[r]) ? sci[e0[r] - ec] : -1;
#pragma unroll
      for (int r = 0; r < R; r++) {
        const bool own = col[r] < a.nb;
        const double sc = own ? s : 1.0;
#pragma unroll
        for (int j = 0; j < BS; j++) xb[r][j] = 0.0;
        if (col[r] >= 0) {
          if (!own && !waited) {
            for (int n = 0; n < a.nwait; n++) p2p_wait(p2p_my_flag(a.P, 0, a.nb_rank[n]), a.hseq, a.P.err);
            waited = true;
          }
          const double *xp = own ? a.x + (size_t)col[r] * BS : a.xg + (size_t)(col[r] - a.nb) * BS;
          if (!own) {
#pragma unroll
            for (int j = 0; j < BS; j++) xb[r][j] = __ldcg(xp + j) * sc;
          } else if (BS == 2) {
            const double2 x2 = *reinterpret_cast<const double2 *>(xp);
            xb[r][0] = x2.x * sc; xb[r][1] = x2.y * sc;
          } else {
#pragma unroll
            for (int j = 0; j < BS; j++) xb[r][j] = xp[j] * sc;
          }
        }
      }
      more = false;
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (col[r] >= 0) {
          const double *vp = sva + (size_t)(e0[r] - ev) * B2;
          double v[B2];
          if (BS == 2) {
            const double2 p = *reinterpret_cast<const double2 *>(vp);
            const double2 q = *reinterpret_cast<const double2 *>(vp + 2);
            v[0] = p.x; v[1] = p.y; v[2] = q.x; v[3] = q.y;
          } else {
#pragma unroll
            for (int q = 0; q < B2; q++) v[q] = vp[q];
          }
#pragma unroll
          for (int j = 0; j < BS; j++)
#pragma unroll
            for (int i = 0; i < BS; i++) acc[r][i] += v[j * BS + i] * xb[r][j];
        }
        e0[r] += 8;
        more = more || (e0[r] < e1[r]);
      }
    }
    // this warp is done with the ring slot
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[st]);
    if (++st == SPMV_TMA_STAGES) {
      st = 0;
      parity ^= 1u;
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
      for (int off = 4; off > 0; off >>= 1)
#pragma unroll
        for (int i = 0; i < BS; i++) acc[r][i] += __shfl_down_sync(0xffffffffu, acc[r][i], off, 8);
      const int row = r0 + grp + r * NG;
      if (row < a.nb && lane == 0) {
        if (BS == 2) *reinterpret_cast<double2 *>(a.y + (size_t)row * 2) = make_double2(acc[r][0], acc[r][1]);
        else {
#pragma unroll
          for (int i = 0; i < BS; i++) a.y[(size_t)row * BS + i] = acc[r][i];
        }
        if (a.xn) {
#pragma unroll
          for (int i = 0; i < BS; i++) a.xn[(size_t)row * BS + i] = a.x[(size_t)row * BS + i] * s;
        }
      }
    }
  }
}

int wb_mat_build_tiles(wb_mat *A) {
  cudaFree(A->d_tile_e0);
  A->d_tile_e0 = nullptr;
  A->ntiles = (A->nb + WB_SPMV_TILE - 1) / WB_SPMV_TILE;
  std::vector<int32_t> e0(A->ntiles + 1, 0);
  int cap = 0;
  for (int t = 0; t < A->ntiles; t++) {
    const int r0 = t * WB_SPMV_TILE, r1 = std::min(A->nb, r0 + WB_SPMV_TILE);
    e0[t] = A->h_rowptr[r0];
    cap = std::max(cap, A->h_rowptr[r1] - A->h_rowptr[r0]);
  }
  e0[A->ntiles] = A->nb > 0 ? A->h_rowptr[A->nb] : 0;
  A->tile_cap = cap;
  WB_CUDA(cudaMalloc(&A->d_tile_e0, sizeof(int32_t) * e0.size()));
  WB_CUDA(wb_memcpy_sync(A->d_tile_e0, e0.data(), sizeof(int32_t) * e0.size(), cudaMemcpyHostToDevice));
  return 0;
}

static size_t spmv_tma_smem(const wb_mat *A) {
  const int b2 = A->bs * A->bs;
  const size_t rp = (WB_SPMV_TILE + 4) * 4, ci = (((size_t)A->tile_cap + 8) * 4 + 15) & ~(size_t)15,
               va = ((size_t)A->tile_cap + 4) * b2 * 8;
  const size_t stage = (rp + ci + va + 127) & ~(size_t)127;
  return 128 + SPMV_TMA_STAGES * stage;
}

static int spmv_tma_mode() {
  static int mode = -1;  // WB_SPMV_TMA = 0 plain kernel, 1 TMA ring with 256 consumer threads, 2 with 512
  if (mode < 0) {
    const char *e = getenv("WB_SPMV_TMA");
    mode = e ? atoi(e) : 0;
  }
  return mode;
}
static bool wb_spmv_tma_enabled(const wb_mat *A) {
  return spmv_tma_mode() != 0 && A->d_tile_e0 && A->ntiles > 0 && spmv_tma_smem(A) <= 110 * 1024;
}

static int wb_spmv_tma_launch(wb_mat *A, const SpmvArgs &a) {
  wb_ctx *c = A->ctx;
  SpmvTmaArgs t = {a, A->d_tile_e0, A->ntiles, A->tile_cap};
  const size_t smem = spmv_tma_smem(A);
  const int grid = std::min(A->ntiles, 2 * WB_NUM_SMS);
#define SPMV_TMA(BS)                                                                                         \
  do {                                                                                                       \
    if (spmv_tma_mode() == 2) {                                                                              \
      cudaFuncSetAttribute(k_bsr_spmv_tma<BS, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      k_bsr_spmv_tma<BS, 512><<<grid, 512 + 32, smem, c->stream>>>(t);                                       \
    } else {                                                                                                 \
      cudaFuncSetAttribute(k_bsr_spmv_tma<BS, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      k_bsr_spmv_tma<BS, 256><<<grid, 256 + 32, smem, c->stream>>>(t);                                       \
    }                                                                                                        \
  } while (0)
  switch (A->bs) {
    case 1: SPMV_TMA(1); break;
    case 2: SPMV_TMA(2); break;
    case 3: SPMV_TMA(3); break;
    default: WB_CHECK(false, "wb_mat_mult: block size %d not supported", A->bs);
  }
#undef SPMV_TMA
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

// ================================================================ vector kernels (K7)


__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[32];
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = l < (blockDim.x >> 5) ? sh[l] : 0.0;
    v = warp_sum(v);
  }
  return v;  // valid in thread 0
}

// true in every thread of the LAST CTA of the grid to get here (its view of the other CTAs' global
// writes is complete); resets the counter for the next launch
__device__ __forceinline__ bool last_block(unsigned *counter) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x * gridDim.y - 1);
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

__global__ void k_gmres_update(const GmresUpd u) { gmres_update(u); }
// multi-GPU: |w|^2 = sum over ranks of slot B (fixed rank order, identical on every rank), then the update
__global__ void k_gmres_update_p2p(const GmresUpd u, const WbP2PDev P, int seq) {
  if (*u.done) return;
  if (threadIdx.x < P.nranks) p2p_wait(p2p_my_flag(P, 2, threadIdx.x), seq, P.err);
  __syncwarp();
  if (threadIdx.x == 0) {
    const double *slot = reinterpret_cast<const double *>(P.region[P.rank] + WB_P2P_SLOT_B);
    double s = 0.0;
    for (int r = 0; r < P.nranks; r++) s += __ldcg(&slot[r]);
    u.scal[0] = s;
    gmres_update(u);
  }
}

// Krylov kernels all take the solver's device-side `done` flag and return at once when it is
// set, so the host can enqueue a whole restart cycle between convergence checks.

// out[j] = w . V_j, j < nv <= NVT, in ONE pass over w and the nv basis vectors (VecMDot): every thread
// keeps nv partial sums; 32-byte vector loads; partials are folded warp -> CTA -> last CTA in a fixed
// order, so the result does not depend on scheduling.
// single-GPU fused BiCGStab (defined with the BiCGStab kernels below)
struct BcgsEpi {
  double *sc;
  KspState *st;
  int *done;
  double rtol, atol, dtol;
  int maxit;
};
__device__ void bcgs_epilogue(const BcgsEpi &e, int which);

struct MdotArgs {
  const double *w, *V;
  size_t ldv;
  int n, nv;
  double *part, *out;
  unsigned *counter;
  const int *done;
  WbP2PDev P;  // seq > 0: the local sums are published to every rank's slot A instead of `out`
  int seq;
  int epi;      // single-GPU fused BiCGStab: scalar recurrence to run in the last CTA after the sums (0 = none)
  BcgsEpi be;
};
template <int NVT, bool VEC>
__global__ void __launch_bounds__(256) k_mdot_all(const MdotArgs a) {
  if (a.done && *a.done) return;
  // blockIdx.y selects a group of NVT (<= 8) basis vectors: every CTA streams its slice of w (L2-resident after
  // the first group) against 8 vectors with all 8 loads of an entry in flight before the first use
  const int jbase = blockIdx.y * NVT;
  const int nv = min(NVT, a.nv - jbase);
  const double *V = a.V + (size_t)jbase * a.ldv;
  double acc[NVT];
#pragma unroll
  for (int j = 0; j < NVT; j++) acc[j] = 0.0;
  if (VEC) {
    const int n4 = a.n >> 2;
    // Traversal from the END of the vectors: the multi-axpy that ran just before (previous iteration) went
    // front to back, so the tail of every basis vector is what the 126 MB L2 still holds; the two passes
    // zig-zag and each one starts on the other's most recently used lines.  The basis is loaded with the
    // default L2 policy, the matrix and factor streams are evict-first.
    for (int i = n4 - 1 - (int)(blockIdx.x * blockDim.x + threadIdx.x); i >= 0; i -= (int)(gridDim.x * blockDim.x)) {
      const double4 wi = ld256(a.w + 4 * (size_t)i);
      double4 v[NVT];
#pragma unroll
      for (int g = 0; g < NVT; g++) {
        // vectors past nv re-read the last one (same cache line, no extra traffic); their sums are dropped
        const int j = g < nv ? g : nv - 1;
        v[g] = ld256(V + (size_t)j * a.ldv + 4 * (size_t)i);
      }
#pragma unroll
      for (int g = 0; g < NVT; g++) {
        acc[g] += wi.x * v[g].x;
        acc[g] += wi.y * v[g].y;
        acc[g] += wi.z * v[g].z;
        acc[g] += wi.w * v[g].w;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {  // tail
      const int i = (n4 << 2) + threadIdx.x;
#pragma unroll
      for (int j = 0; j < NVT; j++)
        if (j < nv) acc[j] += a.w[i] * V[(size_t)j * a.ldv + i];
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
      const double wi = a.w[i];
#pragma unroll
      for (int j = 0; j < NVT; j++)
        if (j < nv) acc[j] += wi * V[(size_t)j * a.ldv + i];
    }
  }
  __shared__ double sh[8][NVT];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NVT; j++) {
    const double v = warp_sum(acc[j]);
    if (lane == 0) sh[wid][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < nv) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += sh[q][threadIdx.x];
    a.part[(size_t)(jbase + threadIdx.x) * RED_BLOCKS + blockIdx.x] = s;
  }
  if (last_block(a.counter)) {
    for (int j = wid; j < a.nv; j += 8) {
      double s = 0.0;
      for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(&a.part[(size_t)j * RED_BLOCKS + b]);
      s = warp_sum(s);
      if (a.seq > 0) {
        // all-gather over NVLink: lane r writes this rank's partial into rank r's slot
        s = __shfl_sync(0xffffffffu, s, 0);
        if (lane < a.P.nranks)
          reinterpret_cast<double *>(a.P.region[lane] + WB_P2P_SLOT_A)[a.P.rank * WB_P2P_MAXV + j] = s;
      } else if (lane == 0) {
        a.out[j] = s;
      }
    }
    if (a.seq > 0) {
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x < a.P.nranks)
        p2p_st_release_sys(reinterpret_cast<int *>(a.P.region[threadIdx.x] + wb_p2p_flag_off(1, a.P.rank)), a.seq);
    } else if (a.epi) {
      __syncthreads();
      if (threadIdx.x == 0) bcgs_epilogue(a.be, a.epi);
    }
  }
}

// w += sign * sum_j coef[j] V_j, j < nv <= NVT (VecMAXPY; sequential in j per entry), in one pass; with
// `part` also |w|^2 of the result (VecNorm), and with `upd` the last CTA finishes the Arnoldi step
// (single-GPU: no collective is needed between the norm and the Hessenberg update).
struct MaxpyArgs {
  double *w;
  const double *V;
  size_t ldv;
  const double *coef;
  double sign;
  int n, nv;
  double *part, *out;  // nullable
  unsigned *counter;
  const int *done;
  int with_upd;
  GmresUpd upd;
  WbP2PDev P;
  int seq_coef;      // > 0: coefficients = sum over ranks of slot A (waits for every rank's sequence number)
  double *coef_out;  // reduced coefficients for the Hessenberg update (written by CTA 0)
  int seq_norm;      // > 0: the local |w|^2 is published to every rank's slot B instead of `out`
  int fuse_tail;     // with seq_norm: the last CTA also waits for every rank's norm, runs the Hessenberg update and
                     // pushes the (scaled) boundary entries of w to the neighbours for the next SpMV
  WbHaloPush push;
};
template <int NVT, bool VEC>
__global__ void __launch_bounds__(256, 2) k_maxpy_all(const MaxpyArgs a) {
  if (a.done && *a.done) return;
  __shared__ double cf[KRY_MAXV];
  if (a.seq_coef > 0) {
    if (threadIdx.x < a.P.nranks) p2p_wait(p2p_my_flag(a.P, 1, threadIdx.x), a.seq_coef, a.P.err);
    __syncthreads();
    if (threadIdx.x < KRY_MAXV) {
      double sum = 0.0;
      if (threadIdx.x < a.nv) {
        const double *slot = reinterpret_cast<const double *>(a.P.region[a.P.rank] + WB_P2P_SLOT_A);
        for (int r = 0; r < a.P.nranks; r++) sum += __ldcg(&slot[r * WB_P2P_MAXV + threadIdx.x]);  // fixed rank order
        if (blockIdx.x == 0 && a.coef_out) a.coef_out[threadIdx.x] = sum;
      }
      cf[threadIdx.x] = a.sign * sum;
    }
  } else if (threadIdx.x < KRY_MAXV) {
    cf[threadIdx.x] = threadIdx.x < a.nv ? a.sign * a.coef[threadIdx.x] : 0.0;
  }
  __syncthreads();
  double nrm = 0.0;
  if (VEC) {
    const int n4 = a.n >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      double4 wi = ld256(a.w + 4 * (size_t)i);
      constexpr int G = (NVT % 8 == 0) ? 8 : (NVT < 4 ? NVT : 4);
#pragma unroll
      for (int j0 = 0; j0 < NVT; j0 += G) {
        double4 v[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
          const int j = (j0 + g) < a.nv ? (j0 + g) : a.nv - 1;  // cf is zero past nv
          v[g] = ld256(a.V + (size_t)j * a.ldv + 4 * (size_t)i);
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
          const double cj = cf[j0 + g];
          wi.x += cj * v[g].x;
          wi.y += cj * v[g].y;
          wi.z += cj * v[g].z;
          wi.w += cj * v[g].w;
        }
      }
      st256(a.w + 4 * (size_t)i, wi);
      nrm += wi.x * wi.x;
      nrm += wi.y * wi.y;
      nrm += wi.z * wi.z;
      nrm += wi.w * wi.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {  // tail
      const int i = (n4 << 2) + threadIdx.x;
      double wi = a.w[i];
      for (int j = 0; j < a.nv; j++) wi += cf[j] * a.V[(size_t)j * a.ldv + i];
      a.w[i] = wi;
      nrm += wi * wi;
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
      double wi = a.w[i];
#pragma unroll
      for (int j = 0; j < NVT; j++)
        if (j < a.nv) wi += cf[j] * a.V[(size_t)j * a.ldv + i];
      a.w[i] = wi;
      nrm += wi * wi;
    }
  }
  if (!a.part) return;
  const double s = block_sum(nrm);
  if (threadIdx.x == 0) a.part[blockIdx.x] = s;
  if (last_block(a.counter)) {
    __shared__ double s_t;
    if (threadIdx.x < 32) {
      double t = 0.0;
      for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) t += __ldcg(&a.part[b]);
      t = warp_sum(t);
      if (threadIdx.x == 0) s_t = t;
    }
    __syncthreads();
    if (a.seq_norm > 0) {
      if (threadIdx.x < a.P.nranks)
        reinterpret_cast<double *>(a.P.region[threadIdx.x] + WB_P2P_SLOT_B)[a.P.rank] = s_t;
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x < a.P.nranks)
        p2p_st_release_sys(reinterpret_cast<int *>(a.P.region[threadIdx.x] + wb_p2p_flag_off(2, a.P.rank)), a.seq_norm);
      if (a.fuse_tail) {
        // |w|^2 over all ranks (fixed rank order), Hessenberg / convergence update, then the halo of the next SpMV
        if (threadIdx.x < a.P.nranks) p2p_wait(p2p_my_flag(a.P, 2, threadIdx.x), a.seq_norm, a.P.err);
        __syncthreads();
        if (threadIdx.x == 0) {
          const double *slot = reinterpret_cast<const double *>(a.P.region[a.P.rank] + WB_P2P_SLOT_B);
          double sum = 0.0;
          for (int r = 0; r < a.P.nranks; r++) sum += __ldcg(&slot[r]);
          a.upd.scal[0] = sum;
          gmres_update(a.upd);
          __threadfence();
        }
        __syncthreads();
        if (!*reinterpret_cast<volatile int *>(a.upd.done)) {
          const double scale = *reinterpret_cast<volatile double *>(a.upd.scal + 1);
          const WbHaloPush &h = a.push;
          // one CTA pushes the whole boundary: 8 entries per thread in flight (index loads, then the gathers of
          // w at L2, then the peer stores) so the push costs a few L2 round trips, not one per entry
          constexpr int U = 8, WMAX = WB_MAX_NP;
          for (int e0 = threadIdx.x; e0 < h.nsend; e0 += U * blockDim.x) {
            int src[U], dr[U], doff[U];
            double v[U][WMAX];
#pragma unroll
            for (int u = 0; u < U; u++) {
              const int e = e0 + u * blockDim.x;
              src[u] = e < h.nsend ? h.idx[e] : -1;
              dr[u] = e < h.nsend ? h.dst_rank[e] : 0;
              doff[u] = e < h.nsend ? h.dst_off[e] : 0;
            }
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
              for (int k = 0; k < WMAX; k++)
                v[u][k] = (src[u] >= 0 && k < h.width) ? __ldcg(&a.w[(size_t)src[u] * h.width + k]) : 0.0;
#pragma unroll
            for (int u = 0; u < U; u++) {
              if (src[u] >= 0) {
                double *ghost = reinterpret_cast<double *>(a.P.region[dr[u]] + WB_P2P_GHOST) + (size_t)doff[u] * h.width;
#pragma unroll
                for (int k = 0; k < WMAX; k++)
                  if (k < h.width) ghost[k] = v[u][k] * scale;
              }
            }
          }
          __threadfence_system();
          __syncthreads();
          if (threadIdx.x < h.nneigh)
            p2p_st_release_sys(reinterpret_cast<int *>(a.P.region[h.nb_rank[threadIdx.x]] + wb_p2p_flag_off(0, a.P.rank)),
                               h.seq);
        }
      }
    } else if (threadIdx.x == 0) {
      a.out[0] = s_t;
      if (a.with_upd) gmres_update(a.upd);
    }
  }
}

// z = a*x + b*y (+ c*w), scalars from device memory (index into a scalar table) or immediates
struct Lin3 {
  const double *x, *y, *w;
  const double *sa, *sb, *sc;  // device scalars (may be null => use immediates)
  double a, b, c;
};
__global__ void __launch_bounds__(256) k_lin3(double *__restrict__ z, Lin3 q, int n, double *__restrict__ part,
                                              const int *done) {
  if (done && *done) return;
  const double a = q.sa ? q.a * (*q.sa) : q.a, b = q.sb ? q.b * (*q.sb) : q.b, cc = q.sc ? q.c * (*q.sc) : q.c;
  double nrm = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v = 0.0;
    if (q.x) v = a * q.x[i];
    if (q.y) v += b * q.y[i];
    if (q.w) v += cc * q.w[i];
    z[i] = v;
    nrm += v * v;
  }
  if (part) {
    const double s = block_sum(nrm);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
  }
}

// out[j] = sum of the partials of dot j, fixed order; one warp per dot
__global__ void k_reduce_final(const double *__restrict__ part, int nblk, int nd, double *__restrict__ out,
                               const int *done) {
  if (done && *done) return;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), l = threadIdx.x & 31;
  if (j >= nd) return;
  double s = 0.0;
  for (int b = l; b < nblk; b += 32) s += part[(size_t)j * RED_BLOCKS + b];
  s = warp_sum(s);
  if (l == 0) out[j] = s;
}

// ================================================================ GMRES (K7)

// start of a restart cycle: res = sqrt(rr); first cycle fixes rnorm0 and tests convergence
__global__ void k_gmres_begin(double *rs, double *scal, KspState *st, int *done, int first, double rtol,
                              double atol, double dtol) {
  const double res = sqrt(scal[0]);
  st->res = res;
  st->it_inner = 0;
  if (first) {
    st->rnorm0 = res;
    st->its = 0;
    st->reason = 0;
    int reason = 0;
    const double ttol = fmax(rtol * res, atol);
    if (res != res) reason = -9;
    else if (res <= ttol) reason = (res < atol) ? 3 : 2;
    if (!reason && res == 0.0) reason = 3;
    if (reason) {
      st->reason = reason;
      *done = 1;
    }
  }
  scal[1] = res > 0.0 ? 1.0 / res : 1.0;
  rs[0] = res;
}

// one thread: back substitution y = H^-1 rs for the it columns built in this cycle
__global__ void k_gmres_solve_y(const double *H, const double *rs, double *yv, int m, const KspState *st) {
  const int it = st->it_inner;
  for (int k = it - 1; k >= 0; k--) {
    double s = rs[k];
    for (int j = k + 1; j < it; j++) s -= H[(size_t)(m + 1) * j + k] * yv[j];
    yv[k] = s / H[(size_t)(m + 1) * k + k];
  }
  for (int k = it; k < m; k++) yv[k] = 0.0;
}

static std::map<wb_ctx *, KspWork> g_work;

static void free_work(KspWork &w) {
  cudaFree(w.V); cudaFree(w.tmp); cudaFree(w.small); cudaFree(w.part); cudaFree(w.d_st); cudaFree(w.d_done);
  cudaFree(w.d_counter);
  cudaFree(w.d_bar);
  cudaFree(w.d_ll);
  cudaFree(w.d_prof);
  if (w.h_st) cudaFreeHost(w.h_st);
}

int wb_ensure_work(wb_ctx *c, size_t n, int m, KspWork **out) {
  KspWork *wq;
  {
    std::lock_guard<std::mutex> lk(wb_registry_mutex());
    wq = &g_work[c];  // node addresses of a std::map are stable
  }
  KspWork &w = *wq;
  const size_t ld = (n + 31) / 32 * 32;
  if (ld <= w.cap && m <= w.m) {
    w.n = n; w.ld = ld;
    w.wbuf = w.tmp + 2 * w.ld;
  } else {
    const int mcap = std::max(m, w.m);
    const size_t cap = std::max(ld, w.cap);
    free_work(w);
    w = KspWork();
    w.ctx = c; w.n = n; w.m = mcap; w.cap = cap;
    w.ld = ld;
    m = mcap;
    WB_CUDA(cudaMalloc(&w.V, sizeof(double) * cap * (m + 1)));
    WB_CUDA(cudaMalloc(&w.tmp, sizeof(double) * cap * 3));
    w.wbuf = w.tmp + 2 * w.ld;
    WB_CUDA(cudaMalloc(&w.small, sizeof(double) * ((size_t)(m + 1) * m + 6 * (m + 2) + 16)));
    WB_CUDA(cudaMalloc(&w.part, sizeof(double) * RED_BLOCKS * KRY_MAXV));
    WB_CUDA(cudaMalloc(&w.d_st, sizeof(KspState)));
    WB_CUDA(cudaMallocHost(&w.h_st, sizeof(KspState)));
    WB_CUDA(cudaMalloc(&w.d_done, sizeof(int)));
    WB_CUDA(cudaMalloc(&w.d_counter, sizeof(unsigned)));
    WB_CUDA(wb_memset_sync(w.d_counter, 0, sizeof(unsigned)));
    WB_CUDA(cudaMalloc(&w.d_bar, 64 * sizeof(int)));
    WB_CUDA(wb_memset_sync(w.d_bar, 0, 64 * sizeof(int)));
    WB_CUDA(cudaMalloc(&w.d_ll, WB_LL_BYTES));
    WB_CUDA(wb_memset_sync(w.d_ll, 0, WB_LL_BYTES));
    WB_CUDA(cudaMalloc(&w.d_prof, WB_PROF_WORDS * sizeof(unsigned long long)));
    WB_CUDA(wb_memset_sync(w.d_prof, 0, WB_PROF_WORDS * sizeof(unsigned long long)));
  }
  *out = &w;
  return 0;
}

KspWork *wb_find_work(wb_ctx *c) {
  std::lock_guard<std::mutex> lk(wb_registry_mutex());
  auto it = g_work.find(c);
  return it == g_work.end() ? nullptr : &it->second;
}

void wb_linalg_release(wb_ctx *c) {
  std::lock_guard<std::mutex> lk(wb_registry_mutex());
  auto it = g_work.find(c);
  if (it == g_work.end()) return;
  free_work(it->second);
  g_work.erase(it);
}

static int red_blocks(size_t n) { return (int)std::min<size_t>(RED_BLOCKS, (n / 4 + 255) / 256 + 1); }

static bool aligned32(const void *p) { return ((uintptr_t)p & 31) == 0; }
template <int NVT> static void launch_mdot(const MdotArgs &a, int nblk, cudaStream_t s) {
  const int ngroup = (a.nv + NVT - 1) / NVT;
  dim3 grid(std::max(nblk / ngroup, 2 * WB_NUM_SMS), ngroup);
  if (aligned32(a.w) && aligned32(a.V) && (a.ldv & 3) == 0) k_mdot_all<NVT, true><<<grid, 256, 0, s>>>(a);
  else k_mdot_all<NVT, false><<<grid, 256, 0, s>>>(a);
}
template <int NVT> static void launch_maxpy(const MaxpyArgs &a, int nblk, cudaStream_t s) {
  if (aligned32(a.w) && aligned32(a.V) && (a.ldv & 3) == 0) k_maxpy_all<NVT, true><<<nblk, 256, 0, s>>>(a);
  else k_maxpy_all<NVT, false><<<nblk, 256, 0, s>>>(a);
}

// dots[j] = w . V_j for j in [0, nd), summed over ranks.  V_j = V + j*ldv (16-byte aligned vectors).
// With p2p_seq (multi-GPU, NVLink path, nd <= WB_P2P_MAXV) the per-rank sums are all-gathered into every rank's
// slot A by the kernel itself and *p2p_seq is the sequence number the consumer (multi_axpy) waits for; no
// collective is launched and d_out is not written here.
static int multi_dot(KspWork &w, const double *d_w, const double *V, size_t ldv, int nd, double *d_out,
                     const int *done, int *p2p_seq = nullptr, int epi = 0, const BcgsEpi *be = nullptr) {
  wb_ctx *c = w.ctx;
  const int nblk = red_blocks(w.n);
  const bool p2p = p2p_seq && c->p2p.on && nd <= WB_P2P_MAXV;
  if (p2p_seq) *p2p_seq = 0;
  for (int j0 = 0; j0 < nd; j0 += KRY_MAXV) {
    const int nv = std::min(KRY_MAXV, nd - j0);
    MdotArgs a = {d_w, V + (size_t)j0 * ldv, ldv, (int)w.n, nv, w.part, d_out + j0, w.d_counter, done, c->p2p.dev, 0, 0, {}};
    if (epi && be && j0 + nv >= nd) {
      a.epi = epi;
      a.be = *be;
    }
    if (p2p) {
      a.seq = ++c->p2p.seq_a;
      *p2p_seq = a.seq;
    }
    if (nv <= 1) launch_mdot<1>(a, nblk, c->stream);
    else if (nv <= 2) launch_mdot<2>(a, nblk, c->stream);
    else if (nv <= 4) launch_mdot<4>(a, nblk, c->stream);
    else launch_mdot<8>(a, nblk, c->stream);
    WB_LAUNCH(c);
  }
  WB_CUDA(cudaGetLastError());
  if (!p2p) WB_TRY(wb_allreduce_sum(c, d_out, nd));
  return 0;
}

// w += sign * sum_j coef[j] V_j ; if d_nrm2: also |w|^2 (summed over ranks); if upd (and one rank): the
// Arnoldi-step update runs in the same launch.  coef_seq > 0: the coefficients are the rank sums of slot A
// published by multi_dot (NVLink path); d_coef then receives the reduced values for the Hessenberg update.
static int multi_axpy(KspWork &w, double *d_w, const double *V, size_t ldv, int nd, const double *d_coef, double sign,
                      double *d_nrm2, const int *done, const GmresUpd *upd, int coef_seq = 0,
                      int *pushed_halo_seq = nullptr, int halo_width = 0) {
  wb_ctx *c = w.ctx;
  const int nblk = red_blocks(w.n);
  const bool fuse_upd = upd && c->nranks <= 1;
  const bool p2p_norm = coef_seq > 0 && d_nrm2 && upd;
  int seq_norm = 0;
  for (int j0 = 0; j0 < nd; j0 += KRY_MAXV) {
    const int nv = std::min(KRY_MAXV, nd - j0);
    const bool last = j0 + nv >= nd;
    MaxpyArgs a;
    memset(&a, 0, sizeof(a));
    a.w = d_w; a.V = V + (size_t)j0 * ldv; a.ldv = ldv; a.coef = d_coef + j0; a.sign = sign; a.n = (int)w.n; a.nv = nv;
    a.part = (last && d_nrm2) ? w.part : nullptr;
    a.out = d_nrm2; a.counter = w.d_counter; a.done = done;
    a.with_upd = (last && fuse_upd) ? 1 : 0;
    if (a.with_upd) a.upd = *upd;
    a.P = c->p2p.dev;
    if (coef_seq > 0) {
      a.seq_coef = coef_seq;
      a.coef_out = const_cast<double *>(d_coef);
      if (p2p_norm && last) {
        a.seq_norm = seq_norm = ++c->p2p.seq_b;
        // Optional (WB_P2P_FUSE_MAX = largest boundary, in cells, to fuse; default 0 = off): one CTA pushing the
        // boundary costs what the two saved launches cost -- measured 244.7 vs 244.2 ms per step on 8 GPUs, and
        // slower on 2 GPUs (10 000-cell faces) -- so the dedicated multi-CTA push kernel stays the default.
        static int fuse_max = -1;
        if (fuse_max < 0) {
          const char *e = getenv("WB_P2P_FUSE_MAX");
          fuse_max = e ? atoi(e) : 0;
        }
        if (pushed_halo_seq && c->halo.nneigh > 0 && c->halo.nneigh <= 256 && c->halo.nsend <= fuse_max) {
          a.fuse_tail = 1;
          a.upd = *upd;
          a.push = wb_p2p_halo_push_args(c, halo_width);
          *pushed_halo_seq = a.push.seq;
        }
      }
    }
    if (nv <= 1) launch_maxpy<1>(a, nblk, c->stream);
    else if (nv <= 2) launch_maxpy<2>(a, nblk, c->stream);
    else if (nv <= 4) launch_maxpy<4>(a, nblk, c->stream);
    else if (nv <= 8) launch_maxpy<8>(a, nblk, c->stream);
    else if (nv <= 12) launch_maxpy<12>(a, nblk, c->stream);
    else if (nv <= 16) launch_maxpy<16>(a, nblk, c->stream);
    else if (nv <= 20) launch_maxpy<20>(a, nblk, c->stream);
    else if (nv <= 24) launch_maxpy<24>(a, nblk, c->stream);
    else if (nv <= 28) launch_maxpy<28>(a, nblk, c->stream);
    else launch_maxpy<32>(a, nblk, c->stream);
    WB_LAUNCH(c);
  }
  WB_CUDA(cudaGetLastError());
  if (seq_norm > 0) {
    if (pushed_halo_seq && *pushed_halo_seq > 0) return 0;  // update + halo push ran in the multi-axpy's last CTA
    k_gmres_update_p2p<<<1, 32, 0, c->stream>>>(*upd, c->p2p.dev, seq_norm);
    WB_LAUNCH(c);
    return 0;
  }
  if (d_nrm2) WB_TRY(wb_allreduce_sum(c, d_nrm2, 1));
  if (upd && !fuse_upd) {
    k_gmres_update<<<1, 1, 0, c->stream>>>(*upd);
    WB_LAUNCH(c);
  }
  return 0;
}

static int lin3(KspWork &w, double *z, const double *x, double a, const double *sa, const double *y, double b,
                const double *sb, const double *v3, double cc, const double *sc, double *d_nrm2, const int *done) {
  wb_ctx *c = w.ctx;
  const int nblk = std::min<int>(RED_BLOCKS, (int)((w.n + 255) / 256));
  Lin3 q = {x, y, v3, sa, sb, sc, a, b, cc};
  k_lin3<<<nblk, 256, 0, c->stream>>>(z, q, (int)w.n, d_nrm2 ? w.part : nullptr, done);
  WB_LAUNCH(c);
  if (d_nrm2) {
    k_reduce_final<<<1, 32, 0, c->stream>>>(w.part, nblk, 1, d_nrm2, done);
    WB_LAUNCH(c);
    WB_TRY(wb_allreduce_sum(c, d_nrm2, 1));
  }
  WB_CUDA(cudaGetLastError());
  return 0;
}


int wb_fetch_state(KspWork &w) {
  wb_ctx *c = w.ctx;
  WB_CUDA(cudaMemcpyAsync(w.h_st, w.d_st, sizeof(KspState), cudaMemcpyDeviceToHost, c->stream));
  if (c->p2p.on) WB_CUDA(cudaMemcpyAsync(c->h_flags + 5, c->d_flags + 5, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  if (c->p2p.on && c->h_flags[5]) {
    wb_set_error("NVLink peer exchange timed out waiting for a peer's sequence number (a rank is missing or out of step)");
    return -4;
  }
  return 0;
}

// how many Krylov iterations are enqueued between host convergence checks in the first restart cycle
// (later cycles are enqueued whole: a solve that needs a second cycle is a long one)
static int g_check_every = 4;
extern "C" int wb_ksp_set_check_every(int k) {
  g_check_every = std::max(1, k);
  return 0;
}

// KSPSolve_GMRES: restarted, classical Gram-Schmidt (no refinement), left preconditioning,
// convergence on the preconditioned residual norm.  Four launches per iteration on one GPU:
//   SpMV (normalises the new basis vector on the fly and stores it), PC apply, fused multi-dot,
//   fused multi-axpy + norm + Hessenberg/Givens update.
static int gmres_dev(wb_mat *A, wb_pc *pc, const wb_ksp_opts *o, const double *d_b, double *d_x, int *its,
                     int *reason, double *rnorm) {
  wb_ctx *c = A->ctx;
  const size_t n = (size_t)A->nb * A->bs;
  const int m = o->restart > 0 ? o->restart : 30;
  KspWork *wp;
  WB_TRY(wb_ensure_work(c, n, m, &wp));
  KspWork &w = *wp;
  const size_t ld = w.ld;
  double *hcol = w.small, *H = hcol + (m + 2), *cs = H + (size_t)(m + 1) * m, *sn = cs + (m + 1),
         *rs = sn + (m + 1), *yv = rs + (m + 2), *scal = yv + (m + 1);
  double *tmp = w.tmp, *wbuf = w.wbuf;
  const GmresUpd upd = {hcol, H, cs, sn, rs, scal, w.d_st, w.d_done, o->rtol, o->atol, o->dtol, m, o->maxit};
  WB_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_done, 0, sizeof(int), c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_st, 0, sizeof(KspState), c->stream));
  bool first = true;
  while (true) {
    // r = M^-1 (b - A x) -> wbuf (unnormalised start vector of the cycle)
    if (first) {
      WB_TRY(wb_pc_apply_dev(pc, d_b, wbuf));
    } else {
      WB_TRY(wb_spmv_launch(A, d_x, tmp));
      WB_TRY(lin3(w, tmp, d_b, 1.0, nullptr, tmp, -1.0, nullptr, nullptr, 0.0, nullptr, nullptr, nullptr));
      WB_TRY(wb_pc_apply_dev(pc, tmp, wbuf));
    }
    WB_TRY(multi_dot(w, wbuf, wbuf, ld, 1, scal, nullptr));
    k_gmres_begin<<<1, 1, 0, c->stream>>>(rs, scal, w.d_st, w.d_done, first ? 1 : 0, o->rtol, o->atol, o->dtol);
    WB_LAUNCH(c);
    if (first) {
      WB_TRY(wb_fetch_state(w));
      if (w.h_st->reason != 0) break;
    }
    int it = 0, halo_seq = 0;
    bool stop = false;
    while (it < m && !stop) {
      const int chunk = first ? std::min(g_check_every, m - it) : m - it;
      for (int q = 0; q < chunk; q++, it++) {
        // the kernels below are no-ops once the device-side done flag is up
        // V_it = wbuf * (1/|wbuf|) stored by the SpMV that also forms tmp = A V_it
        WB_TRY(wb_spmv_fused(A, wbuf, scal + 1, w.V + (size_t)it * ld, tmp, w.d_done, halo_seq));
        halo_seq = 0;
        WB_TRY(wb_pc_apply_dev(pc, tmp, wbuf, w.d_done));
        int coef_seq = 0;
        WB_TRY(multi_dot(w, wbuf, w.V, ld, it + 1, hcol, w.d_done, &coef_seq));
        // multi-GPU NVLink path: the last CTA of the multi-axpy also reduces the norm, updates the Hessenberg
        // matrix and pushes the halo of the next SpMV (which is on wbuf again unless the cycle ends here)
        const bool next_is_wbuf = it + 1 < m && A == &c->J;
        WB_TRY(multi_axpy(w, wbuf, w.V, ld, it + 1, hcol, -1.0, scal, w.d_done, &upd, coef_seq,
                          next_is_wbuf ? &halo_seq : nullptr, A->bs));
      }
      WB_TRY(wb_fetch_state(w));
      if (w.h_st->reason != 0) stop = true;
    }
    first = false;
    // x += sum_j y_j V_j over the columns actually built (the host copy of the state is current)
    const int ncol = w.h_st->it_inner;
    if (ncol > 0) {
      k_gmres_solve_y<<<1, 1, 0, c->stream>>>(H, rs, yv, m, w.d_st);
      WB_LAUNCH(c);
      WB_TRY(multi_axpy(w, d_x, w.V, ld, ncol, yv, 1.0, nullptr, nullptr, nullptr));
    }
    if (w.h_st->reason != 0) break;
  }
  *its = w.h_st->its;
  *reason = w.h_st->reason;
  *rnorm = w.h_st->res;
  return 0;
}

// ================================================================ BiCGStab (K7)

// scalar recurrences of KSPSolve_BCGS on the device.  sc: 0 rho, 1 rhoold, 2 alpha, 3 omegaold,
// 4 beta, 5 d1, 6 omega, 7 dp2, 8 d2, 9 -alpha, 10 -omega, 11 beta*omegaold (negated)
__device__ void bcgs_step(double *sc, int phase, KspState *st, int *done, double rtol, double atol, double dtol,
                          int maxit) {
  if (*done) return;
  if (phase == 0) {  // after rho = (R, RP)
    if (sc[0] == 0.0) {
      st->reason = -5;
      *done = 1;
      return;
    }
    sc[4] = (sc[0] / sc[1]) * (sc[2] / sc[3]);
    sc[11] = -sc[4] * sc[3];
  } else if (phase == 1) {  // after d1 = (V, RP)
    if (sc[5] == 0.0) {
      st->reason = -5;
      *done = 1;
      return;
    }
    sc[2] = sc[0] / sc[5];
    sc[9] = -sc[2];
  } else if (phase == 2) {  // after d1 = (S,T), d2 = (T,T)
    if (sc[8] == 0.0) {
      st->reason = 3;
      st->res = 0.0;
      st->its += 1;
      sc[6] = 0.0;
      sc[10] = 0.0;
      *done = 2;  // x += alpha P still to be applied by the host
      return;
    }
    sc[6] = sc[5] / sc[8];
    sc[10] = -sc[6];
  } else {  // after dp2 = (R,R)
    const double dp = sqrt(sc[7]);
    sc[1] = sc[0];
    sc[3] = sc[6];
    st->res = dp;
    st->its += 1;
    int reason = 0;
    const double ttol = fmax(rtol * st->rnorm0, atol);
    if (dp != dp) reason = -9;
    else if (dp <= ttol) reason = (dp < atol) ? 3 : 2;
    else if (dp >= dtol * st->rnorm0) reason = -4;
    if (!reason && st->its >= maxit) reason = -3;
    if (reason) {
      st->reason = reason;
      *done = 1;
    }
  }
}

__global__ void k_bcgs_step(double *sc, int phase, KspState *st, int *done, double rtol, double atol, double dtol,
                            int maxit) {
  bcgs_step(sc, phase, st, done, rtol, atol, dtol, maxit);
}

__device__ void bcgs_begin(double *sc, KspState *st, int *done, double rtol, double atol) {
  const double dp = sqrt(sc[7]);
  st->res = dp;
  st->rnorm0 = dp;
  st->its = 0;
  st->reason = 0;
  int reason = 0;
  const double ttol = fmax(rtol * dp, atol);
  if (dp != dp) reason = -9;
  else if (dp <= ttol) reason = (dp < atol) ? 3 : 2;
  if (!reason && dp == 0.0) reason = 3;
  if (reason) {
    st->reason = reason;
    *done = 1;
  }
  sc[1] = 1.0;
  sc[2] = 1.0;
  sc[3] = 1.0;
}
__global__ void k_bcgs_begin(double *sc, KspState *st, int *done, double rtol, double atol) {
  bcgs_begin(sc, st, done, rtol, atol);
}

// ---- single-GPU fused BiCGStab: the scalar recurrences run in the last CTA of the reduction that feeds them
// epilogues: 4 = after (R,R) of the start vector: begin + first rho/beta; 1 = after (V,RP): alpha;
// 2 = after (T,S),(T,T) in sc[12..13]: omega
__device__ void bcgs_epilogue(const BcgsEpi &e, int which) {
  double *sc = e.sc;
  if (which == 4) {
    bcgs_begin(sc, e.st, e.done, e.rtol, e.atol);
    sc[0] = sc[7];  // RP = R: rho = (R,RP) = |R|^2
    bcgs_step(sc, 0, e.st, e.done, e.rtol, e.atol, e.dtol, e.maxit);
  } else if (which == 1) {
    bcgs_step(sc, 1, e.st, e.done, e.rtol, e.atol, e.dtol, e.maxit);
  } else if (which == 2) {
    sc[5] = sc[12];
    sc[8] = sc[13];
    bcgs_step(sc, 2, e.st, e.done, e.rtol, e.atol, e.dtol, e.maxit);
  }
}

// X += alpha P + omega S (VecAXPBYPCZ), R = S - omega T (VecWAXPY), |R|^2 and the next rho = (R,RP) in one pass;
// the last CTA then runs the convergence test of this iteration and the beta of the next one
__global__ void __launch_bounds__(256) k_bcgs_update(double *__restrict__ x, double *__restrict__ R,
                                                     const double *__restrict__ P, const double *__restrict__ S,
                                                     const double *__restrict__ T, const double *__restrict__ RP, int n,
                                                     double *__restrict__ part, unsigned *counter, const BcgsEpi e) {
  const int d = *e.done;
  if (d == 1) return;
  const double alpha = e.sc[2], omega = (d == 2) ? 0.0 : e.sc[6], momega = (d == 2) ? 0.0 : e.sc[10];
  double rr = 0.0, rrp = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    x[i] += alpha * P[i] + omega * S[i];
    const double r = S[i] + momega * T[i];
    R[i] = r;
    rr += r * r;
    rrp += r * RP[i];
  }
  const double s1 = block_sum(rr);
  const double s2 = block_sum(rrp);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = s1;
    part[RED_BLOCKS + blockIdx.x] = s2;
  }
  if (last_block(counter)) {
    if (threadIdx.x < 64) {
      const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
      double t = 0.0;
      for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(&part[(size_t)j * RED_BLOCKS + b]);
      t = warp_sum(t);
      if (lane == 0) e.sc[j == 0 ? 7 : 14] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (d == 2) {  // omega undefined ((T,T) = 0): x += alpha P was the whole update; converged (KSPSolve_BCGS)
        *e.done = 1;
      } else {
        bcgs_step(e.sc, 3, e.st, e.done, e.rtol, e.atol, e.dtol, e.maxit);
        if (!*e.done) {
          e.sc[0] = e.sc[14];  // rho of the next iteration
          bcgs_step(e.sc, 0, e.st, e.done, e.rtol, e.atol, e.dtol, e.maxit);
        }
      }
    }
  }
}

// x += alpha*P + omega*S with device scalars
__global__ void __launch_bounds__(256) k_bcgs_xupdate(double *__restrict__ x, const double *__restrict__ P,
                                                      const double *__restrict__ S, const double *sc, int n,
                                                      const int *done, int only_if_done2) {
  if (only_if_done2 ? (*done != 2) : (*done != 0)) return;
  const double alpha = sc[2], omega = only_if_done2 ? 0.0 : sc[6];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    x[i] += alpha * P[i] + omega * S[i];
}

__global__ void k_done_fix(int *done) {
  if (*done == 2) *done = 1;
}

// P = R + beta*(P - omegaold*V)
__global__ void __launch_bounds__(256) k_bcgs_pupdate(double *__restrict__ P, const double *__restrict__ R,
                                                      const double *__restrict__ V, const double *sc, int n,
                                                      const int *done) {
  if (*done) return;
  const double beta = sc[4], omegaold = sc[3];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    P[i] = R[i] + beta * (P[i] - omegaold * V[i]);
}

// single GPU: nine launches per iteration -- P update, SpMV, PC, (V,RP)+alpha, S, SpMV, PC, (T,S),(T,T)+omega,
// X/R update + |R| + next rho + convergence test -- every scalar recurrence in the last CTA of the reduction that
// feeds it (the arithmetic of each recurrence is the k_bcgs_step code of the multi-GPU path)
static int bcgs_dev_fused(wb_mat *A, wb_pc *pc, const wb_ksp_opts *o, const double *d_b, double *d_x, int *its,
                          int *reason, double *rnorm) {
  wb_ctx *c = A->ctx;
  const size_t n = (size_t)A->nb * A->bs;
  KspWork *wp;
  WB_TRY(wb_ensure_work(c, n, 30, &wp));
  KspWork &w = *wp;
  const size_t ld = w.ld;
  double *R = w.V, *RP = R + ld, *P = RP + ld, *V = P + ld, *S = V + ld, *T = S + ld, *tmp = w.tmp;
  double *sc = w.small;
  const int nblk = std::min<int>(RED_BLOCKS, (int)((n + 255) / 256));
  const BcgsEpi be = {sc, w.d_st, w.d_done, o->rtol, o->atol, o->dtol, o->maxit};
  WB_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(P, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(V, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_done, 0, sizeof(int), c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_st, 0, sizeof(KspState), c->stream));
  WB_CUDA(cudaMemsetAsync(sc, 0, sizeof(double) * 16, c->stream));
  WB_TRY(wb_pc_apply_dev(pc, d_b, R));
  WB_TRY(multi_dot(w, R, R, ld, 1, sc + 7, nullptr, nullptr, 4, &be));
  WB_CUDA(cudaMemcpyAsync(RP, R, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
  WB_TRY(wb_fetch_state(w));
  int enq = 0;
  while (w.h_st->reason == 0) {
    for (int q = 0; q < g_check_every && enq < o->maxit; q++, enq++) {
      k_bcgs_pupdate<<<nblk, 256, 0, c->stream>>>(P, R, V, sc, (int)n, w.d_done);
      WB_LAUNCH(c);
      WB_TRY(wb_spmv_fused(A, P, nullptr, nullptr, tmp, w.d_done));
      WB_TRY(wb_pc_apply_dev(pc, tmp, V, w.d_done));
      WB_TRY(multi_dot(w, V, RP, ld, 1, sc + 5, w.d_done, nullptr, 1, &be));
      WB_TRY(lin3(w, S, R, 1.0, nullptr, V, 1.0, sc + 9, nullptr, 0.0, nullptr, nullptr, w.d_done));
      WB_TRY(wb_spmv_fused(A, S, nullptr, nullptr, tmp, w.d_done));
      WB_TRY(wb_pc_apply_dev(pc, tmp, T, w.d_done));
      WB_TRY(multi_dot(w, T, S, ld, 2, sc + 12, w.d_done, nullptr, 2, &be));  // S, T adjacent: (T,S), (T,T) in one pass
      k_bcgs_update<<<nblk, 256, 0, c->stream>>>(d_x, R, P, S, T, RP, (int)n, w.part, w.d_counter, be);
      WB_LAUNCH(c);
    }
    WB_CUDA(cudaGetLastError());
    WB_TRY(wb_fetch_state(w));
    if (enq >= o->maxit && w.h_st->reason == 0) {
      w.h_st->reason = -3;
      break;
    }
  }
  *its = w.h_st->its;
  *reason = w.h_st->reason;
  *rnorm = w.h_st->res;
  return 0;
}

static int bcgs_dev(wb_mat *A, wb_pc *pc, const wb_ksp_opts *o, const double *d_b, double *d_x, int *its,
                    int *reason, double *rnorm) {
  wb_ctx *c = A->ctx;
  if (c->nranks <= 1) return bcgs_dev_fused(A, pc, o, d_b, d_x, its, reason, rnorm);
  const size_t n = (size_t)A->nb * A->bs;
  KspWork *wp;
  WB_TRY(wb_ensure_work(c, n, 30, &wp));
  KspWork &w = *wp;
  // carve the BCGS vectors out of the Krylov basis storage
  const size_t ld = w.ld;
  double *R = w.V, *RP = R + ld, *P = RP + ld, *V = P + ld, *S = V + ld, *T = S + ld, *tmp = w.tmp;
  double *sc = w.small;
  const int nblk = red_blocks(n);
  WB_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(P, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(V, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_done, 0, sizeof(int), c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_st, 0, sizeof(KspState), c->stream));
  WB_CUDA(cudaMemsetAsync(sc, 0, sizeof(double) * 16, c->stream));
  WB_TRY(wb_pc_apply_dev(pc, d_b, R));
  WB_TRY(multi_dot(w, R, R, ld, 1, sc + 7, nullptr));
  k_bcgs_begin<<<1, 1, 0, c->stream>>>(sc, w.d_st, w.d_done, o->rtol, o->atol);
  WB_LAUNCH(c);
  WB_CUDA(cudaMemcpyAsync(RP, R, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
  WB_TRY(wb_fetch_state(w));
  int enq = 0;
  while (w.h_st->reason == 0) {
    for (int q = 0; q < g_check_every && enq < o->maxit; q++, enq++) {
      WB_TRY(multi_dot(w, R, RP, ld, 1, sc + 0, w.d_done));
      k_bcgs_step<<<1, 1, 0, c->stream>>>(sc, 0, w.d_st, w.d_done, o->rtol, o->atol, o->dtol, o->maxit);
      WB_LAUNCH(c);
      k_bcgs_pupdate<<<nblk, 256, 0, c->stream>>>(P, R, V, sc, (int)n, w.d_done);
      WB_LAUNCH(c);
      WB_TRY(wb_spmv_launch(A, P, tmp));
      WB_TRY(wb_pc_apply_dev(pc, tmp, V));
      WB_TRY(multi_dot(w, V, RP, ld, 1, sc + 5, w.d_done));
      k_bcgs_step<<<1, 1, 0, c->stream>>>(sc, 1, w.d_st, w.d_done, o->rtol, o->atol, o->dtol, o->maxit);
      WB_LAUNCH(c);
      WB_TRY(lin3(w, S, R, 1.0, nullptr, V, 1.0, sc + 9, nullptr, 0.0, nullptr, nullptr, w.d_done));
      WB_TRY(wb_spmv_launch(A, S, tmp));
      WB_TRY(wb_pc_apply_dev(pc, tmp, T));
      WB_TRY(multi_dot(w, S, T, ld, 1, sc + 5, w.d_done));
      WB_TRY(multi_dot(w, T, T, ld, 1, sc + 8, w.d_done));
      k_bcgs_step<<<1, 1, 0, c->stream>>>(sc, 2, w.d_st, w.d_done, o->rtol, o->atol, o->dtol, o->maxit);
      WB_LAUNCH(c);
      k_bcgs_xupdate<<<nblk, 256, 0, c->stream>>>(d_x, P, S, sc, (int)n, w.d_done, 1);
      WB_LAUNCH(c);
      k_done_fix<<<1, 1, 0, c->stream>>>(w.d_done);
      WB_LAUNCH(c);
      k_bcgs_xupdate<<<nblk, 256, 0, c->stream>>>(d_x, P, S, sc, (int)n, w.d_done, 0);
      WB_LAUNCH(c);
      WB_TRY(lin3(w, R, S, 1.0, nullptr, T, 1.0, sc + 10, nullptr, 0.0, nullptr, sc + 7, w.d_done));
      k_bcgs_step<<<1, 1, 0, c->stream>>>(sc, 3, w.d_st, w.d_done, o->rtol, o->atol, o->dtol, o->maxit);
      WB_LAUNCH(c);
    }
    WB_CUDA(cudaGetLastError());
    WB_TRY(wb_fetch_state(w));
    if (enq >= o->maxit && w.h_st->reason == 0) {
      w.h_st->reason = -3;
      break;
    }
  }
  *its = w.h_st->its;
  *reason = w.h_st->reason;
  *rnorm = w.h_st->res;
  return 0;
}

// host-visible dot product of two device vectors (summed over ranks); synchronises
int wb_vec_dot_host(wb_ctx *c, const double *d_a, const double *d_b, size_t n, double *out) {
  KspWork *wp;
  {
    bool have = false;
    {
      std::lock_guard<std::mutex> lk(wb_registry_mutex());
      auto it = g_work.find(c);
      have = it != g_work.end() && it->second.n == n;
      if (have) wp = &it->second;
    }
    if (!have) WB_TRY(wb_ensure_work(c, n, 30, &wp));
  }
  KspWork &w = *wp;
  double *sc = w.small;
  WB_TRY(multi_dot(w, d_a, d_b, w.ld, 1, sc, nullptr));
  WB_CUDA(cudaMemcpyAsync(c->h_red, sc, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  *out = c->h_red[0];
  return 0;
}

// z = a*x + b*y on device vectors
int wb_vec_axpby_dev(wb_ctx *c, double *z, double a, const double *x, double b, const double *y, size_t n) {
  Lin3 q = {x, y, nullptr, nullptr, nullptr, nullptr, a, b, 0.0};
  k_lin3<<<red_blocks(n), 256, 0, c->stream>>>(z, q, (int)n, nullptr, nullptr);
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

int wb_ksp_solve_dev(wb_mat *A, wb_pc *pc, const wb_ksp_opts *o, const double *d_b, double *d_x, int *its,
                     int *reason, double *rnorm) {
  WbScopedTimer tm(A->ctx, "ksp_solve");
  // the persistent kernel runs GMRES and BiCGStab over block-Jacobi / ILU(0) sub-domains (wb_fused.cu)
  if (wb_fused_usable(A, pc, o)) return wb_gmres_fused(A, pc, o, d_b, d_x, its, reason, rnorm);
  if (o->type == WB_KSP_BCGS) return bcgs_dev(A, pc, o, d_b, d_x, its, reason, rnorm);
  return gmres_dev(A, pc, o, d_b, d_x, its, reason, rnorm);
}

extern "C" int wb_ksp_solve(wb_mat *A, wb_pc *pc, const wb_ksp_opts *opts, const double *b, double *x, int *its,
                            int *reason, double *rnorm) {
  wb_ctx *c = A->ctx;
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(pc && pc->A == A, "wb_ksp_solve: preconditioner was set up for a different matrix");
  int rc = 0;
  WbStage st(c);
  const size_t n = (size_t)A->nb * A->bs;
  const double *db = st.in(b, n, &rc);
  double *dx = st.out(x, n, &rc);
  if (rc) return rc;
  int its_ = 0, reason_ = 0;
  double rn = 0.0;
  WB_TRY(wb_ksp_solve_dev(A, pc, opts, db, dx, &its_, &reason_, &rn));
  if (its) *its = its_;
  if (reason) *reason = reason_;
  if (rnorm) *rnorm = rn;
  return st.finish();
}
