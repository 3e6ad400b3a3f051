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

#include "wb_common.cuh"

// ================================================================ SpMV (K5)

// Eight lanes per block row: lane l takes blocks rowptr[i]+l, +8, ...  Consecutive rows are
// consecutive in `val`, so a warp streams one contiguous span of the value array (the 78 % of
// the algorithmic bytes) with every sector fully used; x is gathered through L2 (16 MB at
// 1 M cells, resident in the 126 MB L2); a 3-step shuffle folds the eight partial block
// products, in a fixed order.
template <int BS>
__global__ void __launch_bounds__(256) k_bsr_spmv(const int32_t *__restrict__ rowptr,
                                                  const int32_t *__restrict__ colidx,
                                                  const double *__restrict__ val, const double *__restrict__ x,
                                                  double *__restrict__ y, int nb) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = gt >> 3, lane = gt & 7;
  double acc[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) acc[i] = 0.0;
  if (row < nb) {
    const int e1 = rowptr[row + 1];
    for (int e = rowptr[row] + lane; e < e1; e += 8) {
      const int col = colidx[e];
      double xb[BS], v[BS * BS];
      if (BS == 2) {
        const double2 x2 = *reinterpret_cast<const double2 *>(x + (size_t)col * 2);
        xb[0] = x2.x; xb[1] = x2.y;
        const double2 a = __ldcs(reinterpret_cast<const double2 *>(val + (size_t)e * 4));
        const double2 b = __ldcs(reinterpret_cast<const double2 *>(val + (size_t)e * 4) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
      } else {
#pragma unroll
        for (int j = 0; j < BS; j++) xb[j] = x[(size_t)col * BS + j];
#pragma unroll
        for (int q = 0; q < BS * BS; q++) v[q] = __ldcs(val + (size_t)e * BS * BS + q);
      }
#pragma unroll
      for (int j = 0; j < BS; j++)
#pragma unroll
        for (int i = 0; i < BS; i++) acc[i] += v[j * BS + i] * xb[j];
    }
  }
#pragma unroll
  for (int off = 4; off > 0; off >>= 1)
#pragma unroll
    for (int i = 0; i < BS; i++) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], off, 8);
  if (row < nb && lane == 0) {
    if (BS == 2) *reinterpret_cast<double2 *>(y + (size_t)row * 2) = make_double2(acc[0], acc[1]);
    else {
#pragma unroll
      for (int i = 0; i < BS; i++) y[(size_t)row * BS + i] = acc[i];
    }
  }
}

// device pointers; x has nb*bs owned entries.  With ghost columns (multi-GPU) x is copied into
// the matrix's local vector and the ghost entries are filled by the halo exchange
// (MatMult_MPIBAIJ's VecScatter).
int wb_spmv_launch(wb_mat *A, const double *d_x, double *d_y) {
  wb_ctx *c = A->ctx;
  const double *xin = d_x;
  if (A->ncolb > A->nb && c->nranks > 1) {
    WB_CUDA(cudaMemcpyAsync(A->d_xloc, d_x, sizeof(double) * (size_t)A->nb * A->bs, cudaMemcpyDeviceToDevice,
                            c->stream));
    WB_TRY(wb_halo_exchange(c, A->d_xloc, A->bs));
    xin = A->d_xloc;
  }
  const int grid = wb_grid((size_t)A->nb * 8, 256);
  switch (A->bs) {
    case 1: k_bsr_spmv<1><<<grid, 256, 0, c->stream>>>(A->d_rowptr, A->d_colidx, A->d_val, xin, d_y, A->nb); break;
    case 2: k_bsr_spmv<2><<<grid, 256, 0, c->stream>>>(A->d_rowptr, A->d_colidx, A->d_val, xin, d_y, A->nb); break;
    case 3: k_bsr_spmv<3><<<grid, 256, 0, c->stream>>>(A->d_rowptr, A->d_colidx, A->d_val, xin, d_y, A->nb); break;
    default: WB_CHECK(false, "wb_mat_mult: block size %d not supported", A->bs);
  }
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
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
  WB_CUDA(cudaMemcpy(A->h_rowptr.data(), rowptr, sizeof(int32_t) * (nb + 1), cudaMemcpyDefault));
  WB_CUDA(cudaMemcpy(A->h_colidx.data(), colidx, sizeof(int32_t) * nnzb, cudaMemcpyDefault));
  WB_CUDA(cudaMalloc(&A->d_rowptr, sizeof(int32_t) * (nb + 1)));
  WB_CUDA(cudaMalloc(&A->d_colidx, sizeof(int32_t) * std::max(nnzb, 1)));
  WB_CUDA(cudaMalloc(&A->d_val, sizeof(double) * std::max<size_t>((size_t)nnzb * bs * bs, 1)));
  WB_CUDA(cudaMalloc(&A->d_xloc, sizeof(double) * (size_t)ncolb * bs));
  WB_CUDA(cudaMemset(A->d_xloc, 0, sizeof(double) * (size_t)ncolb * bs));
  WB_CUDA(cudaMemcpy(A->d_rowptr, A->h_rowptr.data(), sizeof(int32_t) * (nb + 1), cudaMemcpyHostToDevice));
  WB_CUDA(cudaMemcpy(A->d_colidx, A->h_colidx.data(), sizeof(int32_t) * nnzb, cudaMemcpyHostToDevice));
  if (vals) WB_CUDA(cudaMemcpy(A->d_val, vals, sizeof(double) * (size_t)nnzb * bs * bs, cudaMemcpyDefault));
  else WB_CUDA(cudaMemset(A->d_val, 0, sizeof(double) * (size_t)nnzb * bs * bs));
  *out = A;
  return 0;
}

extern "C" int wb_mat_set_values(wb_mat *A, const double *vals) {
  WB_CUDA(cudaSetDevice(A->ctx->device));
  WB_CUDA(cudaMemcpyAsync(A->d_val, vals, sizeof(double) * (size_t)A->nnzb * A->bs * A->bs, cudaMemcpyDefault,
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

struct wb_pc {
  wb_mat *A = nullptr;
  int type = 0, nb = 0, bs = 0, nblocks = 1;
  double *d_dinv = nullptr;  // pbjacobi: inverted diagonal blocks
  // block-Jacobi ILU(0): factor pattern = matrix pattern restricted to each sub-domain
  int nnzb = 0, nsched_f = 0, nsched_b = 0, nlev_f = 0, nlev_b = 0;
  int32_t *d_rowptr = nullptr, *d_colidx = nullptr, *d_diag = nullptr, *d_src = nullptr;
  int32_t *d_sched_f = nullptr, *d_sched_b = nullptr;  // rows in level order, warp-aligned levels, -1 padded
  double *d_val = nullptr;
  int *d_flag = nullptr;    // per-row completion epoch
  int *d_ticket = nullptr;  // CTA ticket counter
  int epoch = 0;
  // sub-domain resident solve (one CTA per block-Jacobi sub-domain, solution kept in shared memory):
  // level-ordered ELL streams of the L and U factors
  bool blocked = false;
  int nblk = 0, max_block_rows = 0, nent = 0, nlvlrow = 0;
  int4 *d_blk = nullptr;        // per block: row0, nrows, lev0 (forward levels first, then backward), nlev_f | nlev_b << 16
  int4 *d_lev = nullptr;        // per level: rbase, n, nk, ebase
  int32_t *d_blk_rows = nullptr;  // global row of each block-local row
  int32_t *d_lvl_row = nullptr;   // block-local row of each level slot
  int32_t *d_ent_col = nullptr;   // block-local column of each entry (-1: padding)
  int32_t *d_ent_src = nullptr;   // index into d_val of each entry (-1: padding)
  int32_t *d_dinv_src = nullptr;  // index into d_val of the inverted diagonal of each backward level slot
  double *d_ent_val = nullptr, *d_lvl_dinv = nullptr;
};

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


template <class T> static int upload(T **p, const std::vector<T> &v) {
  WB_CUDA(cudaMalloc((void **)p, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) WB_CUDA(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// ---- sub-domain resident ILU(0) solve ---------------------------------------------------
// One CTA per block-Jacobi sub-domain.  The sub-domain's part of the solution lives in shared
// memory for both sweeps, so the only HBM traffic is one streaming read of the factors (stored
// level by level, ELL inside a level: thread r of a level reads consecutive 32-byte blocks) plus
// r in and z out.  Levels are separated by __syncthreads; rows inside a level are independent.
// Per row the blocks are applied in ascending column order, i.e. the arithmetic of the sequential
// MatSolve_SeqBAIJ_N_NaturalOrdering restricted to the sub-domain.
__global__ void k_ilu_repack(const double *__restrict__ fac, const int32_t *__restrict__ ent_src, int nent,
                             const int32_t *__restrict__ dinv_src, int ndinv, int b2, double *__restrict__ ent_val,
                             double *__restrict__ lvl_dinv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nent * b2) {
    const int e = i / b2, q = i - e * b2;
    const int sidx = ent_src[e];
    ent_val[i] = sidx >= 0 ? fac[(size_t)sidx * b2 + q] : 0.0;
  }
  if (i < ndinv * b2) {
    const int e = i / b2, q = i - e * b2;
    const int sidx = dinv_src[e];
    lvl_dinv[i] = sidx >= 0 ? fac[(size_t)sidx * b2 + q] : 0.0;
  }
}

#define WB_ILU_MAXLEV 384
template <int BS>
__global__ void __launch_bounds__(128) k_ilu0_block_solve(const int4 *__restrict__ blk, const int4 *__restrict__ lev,
                                                          const int32_t *__restrict__ blk_rows,
                                                          const int32_t *__restrict__ lvl_row,
                                                          const int32_t *__restrict__ ent_col,
                                                          const double *__restrict__ ent_val,
                                                          const double *__restrict__ lvl_dinv,
                                                          const double *__restrict__ r, double *__restrict__ z) {
  constexpr int B2 = BS * BS;
  extern __shared__ double zs[];
  __shared__ int4 slev[WB_ILU_MAXLEV];
  const int4 d = blk[blockIdx.x];
  const int row0 = d.x, lev0 = d.z, nlf = d.w & 0xffff, nlb = (d.w >> 16) & 0xffff;
  const int nl = nlf + nlb;
  const bool cached = nl <= WB_ILU_MAXLEV;
  if (cached)
    for (int l = threadIdx.x; l < nl; l += blockDim.x) slev[l] = lev[lev0 + l];
  __syncthreads();
  for (int l = 0; l < nl; l++) {
    const int4 L = cached ? slev[l] : lev[lev0 + l];
    const bool fwd = l < nlf;
    for (int rr = threadIdx.x; rr < L.y; rr += blockDim.x) {
      const int li = lvl_row[L.x + rr];
      const int grow = blk_rows[row0 + li];
      double s[BS];
      if (fwd) {
#pragma unroll
        for (int i = 0; i < BS; i++) s[i] = r[(size_t)grow * BS + i];
      } else {
#pragma unroll
        for (int i = 0; i < BS; i++) s[i] = zs[li * BS + i];
      }
      for (int k = 0; k < L.z; k++) {
        const size_t e = (size_t)L.w + (size_t)k * L.y + rr;
        const int col = ent_col[e];
        if (col >= 0) {
          double v[B2];
          if (BS == 2) {
            const double2 a = __ldcs(reinterpret_cast<const double2 *>(ent_val + e * 4));
            const double2 b = __ldcs(reinterpret_cast<const double2 *>(ent_val + e * 4) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
          } else {
#pragma unroll
            for (int q = 0; q < B2; q++) v[q] = __ldcs(ent_val + e * B2 + q);
          }
#pragma unroll
          for (int j = 0; j < BS; j++) {
            const double xj = zs[col * BS + j];
#pragma unroll
            for (int i = 0; i < BS; i++) s[i] -= v[j * BS + i] * xj;
          }
        }
      }
      if (fwd) {
#pragma unroll
        for (int i = 0; i < BS; i++) zs[li * BS + i] = s[i];
      } else {
        double t[BS];
        const double *di = lvl_dinv + ((size_t)L.x + rr) * B2;
#pragma unroll
        for (int i = 0; i < BS; i++) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < BS; j++) acc += di[j * BS + i] * s[j];
          t[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < BS; i++) {
          zs[li * BS + i] = t[i];
          z[(size_t)grow * BS + i] = t[i];
        }
      }
    }
    __syncthreads();
  }
}

// host: level-ordered ELL streams of every sub-domain (symbolic, once per pattern)
static int build_block_streams(wb_pc *pc, const std::vector<int32_t> &blk_of, const std::vector<int32_t> &rowptr,
                               const std::vector<int32_t> &colidx, const std::vector<int32_t> &diag) {
  const int nb = pc->nb;
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
  std::vector<int32_t> lvl_row, ent_col, ent_src, dinv_src;
  std::vector<std::vector<int32_t>> rows_of_level;
  for (int b = 0; b < nblk; b++) {
    const int r0 = bcount[b], nr = bcount[b + 1] - bcount[b];
    int nlev[2] = {0, 0};
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
        const std::vector<int32_t> &rows = rows_of_level[l];
        const int n = (int)rows.size();
        int nk = 0;
        for (int row : rows)
          nk = std::max(nk, pass == 0 ? diag[row] - rowptr[row] : rowptr[row + 1] - diag[row] - 1);
        int4 L;
        L.x = (int)lvl_row.size();
        L.y = n;
        L.z = nk;
        L.w = (int)ent_col.size();
        lev.push_back(L);
        for (int row : rows) {
          lvl_row.push_back(local[row]);
          dinv_src.push_back(pass == 1 ? diag[row] : -1);
        }
        const size_t e0 = ent_col.size();
        ent_col.resize(e0 + (size_t)nk * n, -1);
        ent_src.resize(e0 + (size_t)nk * n, -1);
        for (int q = 0; q < n; q++) {
          const int row = rows[q];
          const int k0 = pass == 0 ? rowptr[row] : diag[row] + 1, k1 = pass == 0 ? diag[row] : rowptr[row + 1];
          for (int k = k0; k < k1; k++) {
            ent_col[e0 + (size_t)(k - k0) * n + q] = local[colidx[k]];
            ent_src[e0 + (size_t)(k - k0) * n + q] = k;
          }
        }
      }
      nlev[pass] = nl;
    }
    WB_CHECK(nlev[0] < 65536 && nlev[1] < 65536, "wb_pc_setup: sub-domain with too many levels");
    blk[b].x = r0;
    blk[b].y = nr;
    blk[b].z = lev0;
    blk[b].w = nlev[0] | (nlev[1] << 16);
  }
  pc->nblk = nblk;
  pc->max_block_rows = maxrows;
  pc->nent = (int)ent_col.size();
  pc->nlvlrow = (int)lvl_row.size();
  const int b2 = pc->bs * pc->bs;
  WB_TRY(upload(&pc->d_blk, blk));
  WB_TRY(upload(&pc->d_lev, lev));
  WB_TRY(upload(&pc->d_blk_rows, blk_rows));
  WB_TRY(upload(&pc->d_lvl_row, lvl_row));
  WB_TRY(upload(&pc->d_ent_col, ent_col));
  WB_TRY(upload(&pc->d_ent_src, ent_src));
  WB_TRY(upload(&pc->d_dinv_src, dinv_src));
  WB_CUDA(cudaMalloc(&pc->d_ent_val, sizeof(double) * std::max<size_t>((size_t)pc->nent * b2, 1)));
  WB_CUDA(cudaMalloc(&pc->d_lvl_dinv, sizeof(double) * std::max<size_t>((size_t)pc->nlvlrow * b2, 1)));
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


static int pc_numeric(wb_pc *pc) {
  wb_mat *A = pc->A;
  wb_ctx *c = A->ctx;
  const int nb = pc->nb;
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
      const size_t nthr = (size_t)std::max(pc->nent, pc->nlvlrow) * bs2;
      k_ilu_repack<<<wb_grid(nthr, 256), 256, 0, c->stream>>>(pc->d_val, pc->d_ent_src, pc->nent, pc->d_dinv_src,
                                                            pc->nlvlrow, bs2, pc->d_ent_val, pc->d_lvl_dinv);
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
                  pc->d_val, pc->d_flag, pc->d_ticket, pc->d_blk, pc->d_lev, pc->d_blk_rows, pc->d_lvl_row,
                  pc->d_ent_col, pc->d_ent_src, pc->d_dinv_src, pc->d_ent_val, pc->d_lvl_dinv};
  for (void *p : ptrs) cudaFree(p);
  delete pc;
  return 0;
}

// PCSetUp: symbolic part on the host (pattern restriction + level schedule), numeric part on the GPU
extern "C" int wb_pc_setup(wb_mat *A, int type, int nblocks, const int32_t *block_of_row, wb_pc **out) {
  wb_ctx *c = A->ctx;
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(type >= WB_PC_NONE && type <= WB_PC_BJACOBI_ILU0, "wb_pc_setup: unknown type %d", type);
  wb_pc *pc = new wb_pc();
  pc->A = A; pc->type = type; pc->nb = A->nb; pc->bs = A->bs; pc->nblocks = std::max(nblocks, 1);
  const int nb = A->nb, bs2 = A->bs * A->bs;
  if (type == WB_PC_PBJACOBI) {
    WB_CUDA(cudaMalloc(&pc->d_dinv, sizeof(double) * (size_t)nb * bs2));
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
    WB_CUDA(cudaMemset(pc->d_flag, 0, sizeof(int) * nb));
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

int wb_pc_apply_dev(wb_pc *pc, const double *d_r, double *d_z) {
  wb_ctx *c = pc->A->ctx;
  const int nb = pc->nb;
  if (pc->type == WB_PC_NONE) {
    if (d_r != d_z)
      WB_CUDA(cudaMemcpyAsync(d_z, d_r, sizeof(double) * (size_t)nb * pc->bs, cudaMemcpyDeviceToDevice, c->stream));
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
    const size_t smem = (size_t)pc->max_block_rows * pc->bs * sizeof(double);
#define BSOLVE(BS)                                                                                                 \
  do {                                                                                                             \
    if (smem > 48 * 1024)                                                                                          \
      cudaFuncSetAttribute(k_ilu0_block_solve<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    k_ilu0_block_solve<BS><<<pc->nblk, 128, smem, c->stream>>>(pc->d_blk, pc->d_lev, pc->d_blk_rows, pc->d_lvl_row, \
                                                               pc->d_ent_col, pc->d_ent_val, pc->d_lvl_dinv, d_r, d_z); \
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
    WB_TRY(wb_pc_apply_dev(pc, dr, dz));
  }
  return st.finish();
}

// ================================================================ vector kernels (K7)

#define RED_BLOCKS (4 * WB_NUM_SMS)

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[32];
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = l < (blockDim.x >> 5) ? sh[l] : 0.0;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  }
  return v;  // valid in thread 0
}

// Krylov kernels all take the solver's device-side `done` flag and return at once when it is
// set, so the host can enqueue several iterations between convergence checks.

// part[j*RED_BLOCKS + blk] = partial (w . V_j) for up to 8 vectors per launch
template <int NV>
__global__ void __launch_bounds__(256) k_mdot(const double *__restrict__ w, const double *__restrict__ V, size_t ldv,
                                              int n, double *__restrict__ part, const int *done) {
  if (done && *done) return;
  double acc[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) acc[j] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double wi = w[i];
#pragma unroll
    for (int j = 0; j < NV; j++) acc[j] += wi * V[(size_t)j * ldv + i];
  }
#pragma unroll
  for (int j = 0; j < NV; j++) {
    const double s = block_sum(acc[j]);
    if (threadIdx.x == 0) part[(size_t)j * RED_BLOCKS + blockIdx.x] = s;
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
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if (l == 0) out[j] = s;
}

// w += sign * sum_j coef[j] V_j (sequential in j, as repeated VecAXPY); optional partial |w|^2
template <int NV>
__global__ void __launch_bounds__(256) k_maxpy(double *__restrict__ w, const double *__restrict__ V, size_t ldv,
                                               const double *__restrict__ coef, double sign, int n,
                                               double *__restrict__ part, const int *done) {
  if (done && *done) return;
  double cf[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) cf[j] = sign * coef[j];
  double nrm = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double wi = w[i];
#pragma unroll
    for (int j = 0; j < NV; j++) wi += cf[j] * V[(size_t)j * ldv + i];
    w[i] = wi;
    nrm += wi * wi;
  }
  if (part) {
    const double s = block_sum(nrm);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
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

// ================================================================ GMRES (K7)

struct GmresDev {
  // layout of the small device state block (doubles)
  // [0..m]   hcol (dots of the current column, then rotated)
  // then H (m+1)*m, cs m+1, sn m+1, rs m+2, yv m+1, scal[8]: 0 tt2, 1 scale, 2 res, 3 rnorm0, 4 res0sq
};

struct KspState {
  double res, rnorm0;
  int its, reason, it_inner, pad;
};

// one thread: finish Arnoldi column `it`, update the Givens QR and the convergence state
// (KSPGMRESUpdateHessenberg + KSPConvergedDefault)
__global__ void k_gmres_update(double *hcol, double *H, double *cs, double *sn, double *rs, double *scal, int m,
                               KspState *st, int *done, double rtol, double atol, double dtol, int maxit) {
  if (*done) return;
  const int it = st->it_inner;
  const double tt = sqrt(scal[0]);
  hcol[it + 1] = tt;
  const bool happy = (tt < 1.e-30 * fmax(st->res, 1e-300)) || tt == 0.0;
  scal[1] = happy ? 1.0 : 1.0 / tt;
  double *Hc = H + (size_t)(m + 1) * it;
  for (int j = 0; j <= it + 1; j++) Hc[j] = hcol[j];
  for (int j = 0; j < it; j++) {
    const double t1 = Hc[j], t2 = Hc[j + 1];
    Hc[j] = cs[j] * t1 + sn[j] * t2;
    Hc[j + 1] = -sn[j] * t1 + cs[j] * t2;
  }
  const double hh = Hc[it], hp = Hc[it + 1];
  const double den = sqrt(hh * hh + hp * hp);
  if (den == 0.0) {
    st->reason = -5;  // KSP_DIVERGED_BREAKDOWN
    *done = 1;
    return;
  }
  cs[it] = hh / den;
  sn[it] = hp / den;
  rs[it + 1] = -sn[it] * rs[it];
  rs[it] = cs[it] * rs[it];
  Hc[it] = cs[it] * hh + sn[it] * hp;
  Hc[it + 1] = 0.0;
  const double res = fabs(rs[it + 1]);
  st->res = res;
  st->it_inner = it + 1;
  st->its += 1;
  int reason = 0;
  const double ttol = fmax(rtol * st->rnorm0, atol);
  if (res != res) reason = -9;
  else if (res <= ttol) reason = (res < atol) ? 3 : 2;
  else if (res >= dtol * st->rnorm0) reason = -4;
  if (!reason && happy) reason = 5;
  if (!reason && st->its >= maxit) reason = -3;
  if (reason) {
    st->reason = reason;
    *done = 1;
  }
}

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

struct KspWork {
  wb_ctx *ctx = nullptr;
  size_t n = 0;
  int m = 0;
  double *V = nullptr, *tmp = nullptr, *small = nullptr, *part = nullptr;
  KspState *d_st = nullptr, *h_st = nullptr;
  int *d_done = nullptr;
  double *bc[8] = {nullptr};  // BCGS vectors
};
static std::map<wb_ctx *, KspWork> g_work;

static int ensure_work(wb_ctx *c, size_t n, int m, KspWork **out) {
  KspWork &w = g_work[c];
  if (w.n != n || w.m < m) {
    cudaFree(w.V); cudaFree(w.tmp); cudaFree(w.small); cudaFree(w.part); cudaFree(w.d_st); cudaFree(w.d_done);
    if (w.h_st) cudaFreeHost(w.h_st);
    w = KspWork();
    w.ctx = c; w.n = n; w.m = m;
    WB_CUDA(cudaMalloc(&w.V, sizeof(double) * n * (m + 1)));
    WB_CUDA(cudaMalloc(&w.tmp, sizeof(double) * n * 2));
    WB_CUDA(cudaMalloc(&w.small, sizeof(double) * ((size_t)(m + 1) * m + 6 * (m + 2) + 16)));
    WB_CUDA(cudaMalloc(&w.part, sizeof(double) * RED_BLOCKS * 8));
    WB_CUDA(cudaMalloc(&w.d_st, sizeof(KspState)));
    WB_CUDA(cudaMallocHost(&w.h_st, sizeof(KspState)));
    WB_CUDA(cudaMalloc(&w.d_done, sizeof(int)));
  }
  *out = &w;
  return 0;
}

void wb_linalg_release(wb_ctx *c) {
  auto it = g_work.find(c);
  if (it == g_work.end()) return;
  KspWork &w = it->second;
  cudaFree(w.V); cudaFree(w.tmp); cudaFree(w.small); cudaFree(w.part); cudaFree(w.d_st); cudaFree(w.d_done);
  if (w.h_st) cudaFreeHost(w.h_st);
  g_work.erase(it);
}

static int red_blocks(size_t n) { return (int)std::min<size_t>(RED_BLOCKS, (n + 255) / 256); }

// dots[j] = w . V_j for j in [0, nd), summed over ranks
static int multi_dot(KspWork &w, const double *d_w, const double *V, int nd, double *d_out, const int *done) {
  wb_ctx *c = w.ctx;
  const int n = (int)w.n, nblk = red_blocks(w.n);
  for (int j0 = 0; j0 < nd; j0 += 8) {
    const int nv = std::min(8, nd - j0);
    const double *Vj = V + (size_t)j0 * w.n;
    switch (nv) {
      case 1: k_mdot<1><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, n, w.part, done); break;
      case 2: k_mdot<2><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, n, w.part, done); break;
      case 3: k_mdot<3><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, n, w.part, done); break;
      case 4: k_mdot<4><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, n, w.part, done); break;
      case 5: k_mdot<5><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, n, w.part, done); break;
      case 6: k_mdot<6><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, n, w.part, done); break;
      case 7: k_mdot<7><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, n, w.part, done); break;
      default: k_mdot<8><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, n, w.part, done); break;
    }
    WB_LAUNCH(c);
    k_reduce_final<<<1, 256, 0, c->stream>>>(w.part, nblk, nv, d_out + j0, done);
    WB_LAUNCH(c);
  }
  WB_CUDA(cudaGetLastError());
  WB_TRY(wb_allreduce_sum(c, d_out, nd));
  return 0;
}

// w += sign * sum_j coef[j] V_j ; if d_nrm2: also |w|^2 (summed over ranks)
static int multi_axpy(KspWork &w, double *d_w, const double *V, int nd, const double *d_coef, double sign,
                      double *d_nrm2, const int *done) {
  wb_ctx *c = w.ctx;
  const int n = (int)w.n, nblk = red_blocks(w.n);
  for (int j0 = 0; j0 < nd; j0 += 8) {
    const int nv = std::min(8, nd - j0);
    const bool last = j0 + nv >= nd;
    double *part = (last && d_nrm2) ? w.part : nullptr;
    const double *Vj = V + (size_t)j0 * w.n;
    switch (nv) {
      case 1: k_maxpy<1><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, d_coef + j0, sign, n, part, done); break;
      case 2: k_maxpy<2><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, d_coef + j0, sign, n, part, done); break;
      case 3: k_maxpy<3><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, d_coef + j0, sign, n, part, done); break;
      case 4: k_maxpy<4><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, d_coef + j0, sign, n, part, done); break;
      case 5: k_maxpy<5><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, d_coef + j0, sign, n, part, done); break;
      case 6: k_maxpy<6><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, d_coef + j0, sign, n, part, done); break;
      case 7: k_maxpy<7><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, d_coef + j0, sign, n, part, done); break;
      default: k_maxpy<8><<<nblk, 256, 0, c->stream>>>(d_w, Vj, w.n, d_coef + j0, sign, n, part, done); break;
    }
    WB_LAUNCH(c);
  }
  if (d_nrm2) {
    k_reduce_final<<<1, 32, 0, c->stream>>>(w.part, nblk, 1, d_nrm2, done);
    WB_LAUNCH(c);
    WB_TRY(wb_allreduce_sum(c, d_nrm2, 1));
  }
  WB_CUDA(cudaGetLastError());
  return 0;
}

static int lin3(KspWork &w, double *z, const double *x, double a, const double *sa, const double *y, double b,
                const double *sb, const double *v3, double cc, const double *sc, double *d_nrm2, const int *done) {
  wb_ctx *c = w.ctx;
  const int nblk = red_blocks(w.n);
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

int wb_pc_apply_dev(wb_pc *pc, const double *d_r, double *d_z);

static int fetch_state(KspWork &w) {
  wb_ctx *c = w.ctx;
  WB_CUDA(cudaMemcpyAsync(w.h_st, w.d_st, sizeof(KspState), cudaMemcpyDeviceToHost, c->stream));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// how many Krylov iterations are enqueued between host convergence checks
static int g_check_every = 4;
extern "C" int wb_ksp_set_check_every(int k) {
  g_check_every = std::max(1, k);
  return 0;
}

// KSPSolve_GMRES: restarted, classical Gram-Schmidt (no refinement), left preconditioning,
// convergence on the preconditioned residual norm
static int gmres_dev(wb_mat *A, wb_pc *pc, const wb_ksp_opts *o, const double *d_b, double *d_x, int *its,
                     int *reason, double *rnorm) {
  wb_ctx *c = A->ctx;
  const size_t n = (size_t)A->nb * A->bs;
  const int m = o->restart > 0 ? o->restart : 30;
  KspWork *wp;
  WB_TRY(ensure_work(c, n, m, &wp));
  KspWork &w = *wp;
  double *hcol = w.small, *H = hcol + (m + 2), *cs = H + (size_t)(m + 1) * m, *sn = cs + (m + 1),
         *rs = sn + (m + 1), *yv = rs + (m + 2), *scal = yv + (m + 1);
  double *tmp = w.tmp;
  WB_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_done, 0, sizeof(int), c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_st, 0, sizeof(KspState), c->stream));
  bool first = true;
  while (true) {
    // r = M^-1 (b - A x) -> V_0
    if (first) {
      WB_TRY(wb_pc_apply_dev(pc, d_b, w.V));
    } else {
      WB_TRY(wb_spmv_launch(A, d_x, tmp));
      WB_TRY(lin3(w, tmp, d_b, 1.0, nullptr, tmp, -1.0, nullptr, nullptr, 0.0, nullptr, nullptr, nullptr));
      WB_TRY(wb_pc_apply_dev(pc, tmp, w.V));
    }
    WB_TRY(multi_dot(w, w.V, w.V, 1, scal, nullptr));
    k_gmres_begin<<<1, 1, 0, c->stream>>>(rs, scal, w.d_st, w.d_done, first ? 1 : 0, o->rtol, o->atol, o->dtol);
    WB_LAUNCH(c);
    if (first) {
      WB_TRY(fetch_state(w));
      if (w.h_st->reason != 0) break;
    }
    first = false;
    WB_TRY(lin3(w, w.V, w.V, 1.0, scal + 1, nullptr, 0.0, nullptr, nullptr, 0.0, nullptr, nullptr, w.d_done));
    int it = 0;
    bool stop = false;
    while (it < m && !stop) {
      const int chunk = std::min(g_check_every, m - it);
      for (int q = 0; q < chunk; q++, it++) {
        double *vn = w.V + (size_t)(it + 1) * n;
        // the kernels below are no-ops once the device-side done flag is up
        WB_TRY(wb_spmv_launch(A, w.V + (size_t)it * n, tmp));
        WB_TRY(wb_pc_apply_dev(pc, tmp, vn));
        WB_TRY(multi_dot(w, vn, w.V, it + 1, hcol, w.d_done));
        WB_TRY(multi_axpy(w, vn, w.V, it + 1, hcol, -1.0, scal, w.d_done));
        k_gmres_update<<<1, 1, 0, c->stream>>>(hcol, H, cs, sn, rs, scal, m, w.d_st, w.d_done, o->rtol, o->atol,
                                               o->dtol, o->maxit);
        WB_LAUNCH(c);
        WB_TRY(lin3(w, vn, vn, 1.0, scal + 1, nullptr, 0.0, nullptr, nullptr, 0.0, nullptr, nullptr, w.d_done));
      }
      WB_TRY(fetch_state(w));
      if (w.h_st->reason != 0) stop = true;
    }
    // x += sum_j y_j V_j over the columns actually built (the host copy of the state is current)
    const int ncol = w.h_st->it_inner;
    if (ncol > 0) {
      k_gmres_solve_y<<<1, 1, 0, c->stream>>>(H, rs, yv, m, w.d_st);
      WB_LAUNCH(c);
      WB_TRY(multi_axpy(w, d_x, w.V, ncol, yv, 1.0, nullptr, nullptr));
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
__global__ void k_bcgs_step(double *sc, int phase, KspState *st, int *done, double rtol, double atol, double dtol,
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

__global__ void k_bcgs_begin(double *sc, KspState *st, int *done, double rtol, double atol) {
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

static int bcgs_dev(wb_mat *A, wb_pc *pc, const wb_ksp_opts *o, const double *d_b, double *d_x, int *its,
                    int *reason, double *rnorm) {
  wb_ctx *c = A->ctx;
  const size_t n = (size_t)A->nb * A->bs;
  KspWork *wp;
  WB_TRY(ensure_work(c, n, 30, &wp));
  KspWork &w = *wp;
  // carve the BCGS vectors out of the Krylov basis storage
  double *R = w.V, *RP = R + n, *P = RP + n, *V = P + n, *S = V + n, *T = S + n, *tmp = w.tmp;
  double *sc = w.small;
  const int nblk = red_blocks(n);
  WB_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(P, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(V, 0, sizeof(double) * n, c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_done, 0, sizeof(int), c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_st, 0, sizeof(KspState), c->stream));
  WB_CUDA(cudaMemsetAsync(sc, 0, sizeof(double) * 16, c->stream));
  WB_TRY(wb_pc_apply_dev(pc, d_b, R));
  WB_TRY(multi_dot(w, R, R, 1, sc + 7, nullptr));
  k_bcgs_begin<<<1, 1, 0, c->stream>>>(sc, w.d_st, w.d_done, o->rtol, o->atol);
  WB_LAUNCH(c);
  WB_CUDA(cudaMemcpyAsync(RP, R, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
  WB_TRY(fetch_state(w));
  int enq = 0;
  while (w.h_st->reason == 0) {
    for (int q = 0; q < g_check_every && enq < o->maxit; q++, enq++) {
      WB_TRY(multi_dot(w, R, RP, 1, sc + 0, w.d_done));
      k_bcgs_step<<<1, 1, 0, c->stream>>>(sc, 0, w.d_st, w.d_done, o->rtol, o->atol, o->dtol, o->maxit);
      WB_LAUNCH(c);
      k_bcgs_pupdate<<<nblk, 256, 0, c->stream>>>(P, R, V, sc, (int)n, w.d_done);
      WB_LAUNCH(c);
      WB_TRY(wb_spmv_launch(A, P, tmp));
      WB_TRY(wb_pc_apply_dev(pc, tmp, V));
      WB_TRY(multi_dot(w, V, RP, 1, sc + 5, w.d_done));
      k_bcgs_step<<<1, 1, 0, c->stream>>>(sc, 1, w.d_st, w.d_done, o->rtol, o->atol, o->dtol, o->maxit);
      WB_LAUNCH(c);
      WB_TRY(lin3(w, S, R, 1.0, nullptr, V, 1.0, sc + 9, nullptr, 0.0, nullptr, nullptr, w.d_done));
      WB_TRY(wb_spmv_launch(A, S, tmp));
      WB_TRY(wb_pc_apply_dev(pc, tmp, T));
      WB_TRY(multi_dot(w, S, T, 1, sc + 5, w.d_done));
      WB_TRY(multi_dot(w, T, T, 1, sc + 8, w.d_done));
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
    WB_TRY(fetch_state(w));
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
    auto it = g_work.find(c);
    if (it == g_work.end() || it->second.n != n) WB_TRY(ensure_work(c, n, 30, &wp));
    else wp = &it->second;
  }
  KspWork &w = *wp;
  double *sc = w.small;
  WB_TRY(multi_dot(w, d_a, d_b, 1, sc, nullptr));
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
