// wb_linalg.cuh -- declarations shared by the linear-algebra translation units (wb_linalg.cu: SpMV, preconditioners,
// launch-per-operation Krylov solvers; wb_fused.cu: the sub-domain-resident persistent GMRES kernel): NVLink P2P
// primitives, 256-bit accesses, mbarrier / bulk-copy (TMA) wrappers, the plane-layout ILU(0) level kernel, the
// preconditioner object, the GMRES scalar update and the Krylov work space.
#pragma once
#include "wb_common.cuh"

// ================================================================ NVLink P2P helpers

__device__ __forceinline__ int p2p_ld_acquire_sys(const int *p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void p2p_st_release_sys(int *p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spin until the peer's sequence number reaches `seq`; bounded (~15 s) so that a lost peer raises an error flag
// instead of hanging the GPU
__device__ __forceinline__ void p2p_wait(const int *flag, int seq, int *err) {
  const long long t0 = clock64();
  while (p2p_ld_acquire_sys(flag) < seq) {
    if (clock64() - t0 > 30000000000LL) {
      atomicExch(err, 1);
      break;
    }
  }
}
__device__ __forceinline__ const int *p2p_my_flag(const WbP2PDev &P, int kind, int sender) {
  return reinterpret_cast<const int *>(P.region[P.rank] + wb_p2p_flag_off(kind, sender));
}

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256); pointers must be 32-byte aligned
__device__ __forceinline__ double4 ld256(const double *p) {
  double4 r;
  asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double4 ld256_stream(const double *p) {  // evict-first variant (single-use streams)
  double4 r;
  asm("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st256(double *p, const double4 &v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

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
  int nblk = 0, max_block_rows = 0, nrepack = 0;
  int4 *d_blk = nullptr;          // per block: row0, nrows, lev0 (forward levels first, then backward), number of levels
  int4 *d_lev = nullptr;          // per level: word offset into d_stream, bytes, rows, unused
  int32_t *d_blk_rows = nullptr;  // global row of each block-local row
  double *d_stream = nullptr;     // level-ordered factor stream (see "sub-domain resident ILU(0) solve")
  size_t stream_words = 0;
  int4 *d_repack = nullptr;       // (block index into d_val, word offset into d_stream, plane stride, 0) of every factor block
  int stage_words = 0, nstage = 0, desc_words = 0, solve_threads = 128;
  long long *d_trace = nullptr;  // debug timeline of the sub-domain solve (wb_debug_pc_trace)
  struct WbFusedPlan *fused = nullptr;
  // restricted additive Schwarz, overlap 1 (WB_PC_ASM_ILU0): the matrices of the extended sub-domains side by side as
  // one block-diagonal matrix, its block-Jacobi ILU(0), and the maps between the two row numberings
  wb_mat *asm_mat = nullptr;
  wb_pc *asm_inner = nullptr;
  int asm_ne = 0;
  int32_t *d_asm_src = nullptr, *d_asm_row = nullptr, *d_asm_own = nullptr;
  double *d_asm_r = nullptr, *d_asm_z = nullptr;
  // host copies of the sub-domain tables (symbolic), kept for the fused plan
  std::vector<int4> h_blk, h_lev;
  std::vector<int32_t> h_blk_rows;  // sub-domain-resident persistent GMRES (wb_fused.cu); null: not available
};

template <class T> static int upload(T **p, const std::vector<T> &v) {
  WB_CUDA(cudaMalloc((void **)p, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) WB_CUDA(wb_memcpy_sync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// ---- sub-domain resident ILU(0) solve ---------------------------------------------------
// One CTA per block-Jacobi sub-domain.  The sub-domain's part of the solution lives in shared
// memory for both sweeps, so the only HBM traffic is ONE streaming read of the factors plus r in
// and z out.  The factors are stored as a level-ordered stream: for every dependency level of the
// forward sweep, then of the backward sweep, one contiguous 16-byte-aligned record of n rows with
// nk off-diagonal blocks each (ELL, padded with zero blocks that point at a zero slot of the vector):
//     index planes   int4[NI][n]      (local row, col_0, col_1, col_2), (col_3 .. col_6), ...
//     value planes   double[PW]-wide planes [ (nk + bwd) * bs*bs/PW ][n]   L or U blocks, then the
//                    inverted diagonal block (backward sweep only); PW = 2 for even bs*bs
// Thread r of a level reads element r of every plane: conflict-free shared-memory accesses.  A whole
// level is ONE TMA bulk copy (cp.async.bulk, mbarrier-completed) into a shared-memory ring `nstage`
// levels deep: the copy of level l+nstage is in flight while level l is applied, which takes the HBM
// latency off the level-to-level dependency chain.  Rows inside a level are independent; levels are
// separated by __syncthreads.  Per row the blocks are applied in ascending column order, i.e. the
// arithmetic of the sequential MatSolve_SeqBAIJ_N_NaturalOrdering restricted to the sub-domain.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bulk copy global -> shared, completing on an mbarrier; L2 evict-first: the factor stream is read once per
// apply and must not displace the Krylov basis from L2
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int BS> struct IluPlane {
  static constexpr int B2 = BS * BS;
  static constexpr int PW = (B2 % 2 == 0) ? 2 : 1;  // doubles per plane element
  static constexpr int NP = B2 / PW;                // planes per block
};

// ---- level records of the sub-domain resident ILU(0) sweeps (row-major) ----------------------------------------
// One record per dependency level: n rows of `stride` bytes each,
//     int32[4 * ni4]   byte offset into the sub-domain vector of the row's own entry, then of the nk column entries
//                      (padding columns point at the zero slot behind the vector)
//     double[nk * bs*bs]  the row's L (forward) or U (backward) blocks, column-major, ascending column order
//     double[bs*bs]       backward only: the inverted diagonal block
// with the stride padded to 16 (mod 32) bytes: thread r of a level reads everything at constant offsets from ONE
// address (base + r * stride), and the 16-byte accesses of eight consecutive lanes fall into disjoint bank groups
// (stride / 4 = 4 * odd words), i.e. the shared-memory accesses are conflict-free without a plane layout.  The sweep
// is a chain of dependent levels executed by very few warps, so what counts is the number of dependent instructions
// per level, not throughput: no per-plane address arithmetic, indices pre-multiplied.
__host__ __device__ __forceinline__ int ilu_ni4(int nk) { return (nk + 4) / 4; }  // int4 words of the index area
__host__ __device__ __forceinline__ int ilu_row_stride(int bs, int nk, int bwd) {
  const int raw = ilu_ni4(nk) * 16 + (nk + bwd) * bs * bs * 8;
  const int s16 = (raw + 15) & ~15;
  return (s16 & 16) ? s16 : s16 + 16;  // 16 (mod 32)
}

template <int BS> __device__ __forceinline__ void ilu_ld_block(const double *bp, double *v) {
  constexpr int B2 = BS * BS;
  if (B2 % 2 == 0) {
#pragma unroll
    for (int q = 0; q < B2 / 2; q++) {
      const double2 t = *reinterpret_cast<const double2 *>(bp + 2 * q);
      v[2 * q] = t.x;
      v[2 * q + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int q = 0; q < B2; q++) v[q] = bp[q];
  }
}
template <int BS> __device__ __forceinline__ void ilu_ld_vec(const unsigned char *zb, int off, double *x) {
  if (BS == 2) {
    const double2 t = *reinterpret_cast<const double2 *>(zb + off);
    x[0] = t.x;
    x[1] = t.y;
  } else {
#pragma unroll
    for (int j = 0; j < BS; j++) x[j] = reinterpret_cast<const double *>(zb + off)[j];
  }
}

// apply one level record (in shared or global memory) to the sub-domain vector zs.  The common nk == 3 case
// (7-point stencils) is branch-free with every load issued before the first use, so a level costs one shared-memory
// round trip plus a short dependent FMA chain.  Per row the products are those of the sequential
// MatSolve_SeqBAIJ_N_NaturalOrdering restricted to the sub-domain.
template <int BS>
__device__ __forceinline__ void ilu_level(const double *lv, int n, int nk, bool bwd, double *zs, int nthr, int tid) {
  constexpr int B2 = BS * BS;
  const int stride = ilu_row_stride(BS, nk, bwd ? 1 : 0), ib = ilu_ni4(nk) * 16;
  const unsigned char *base = reinterpret_cast<const unsigned char *>(lv);
  unsigned char *zb = reinterpret_cast<unsigned char *>(zs);
  for (int rr = tid; rr < n; rr += nthr) {
    const unsigned char *rp = base + (size_t)rr * stride;
    const int4 i0 = *reinterpret_cast<const int4 *>(rp);
    const double *bp = reinterpret_cast<const double *>(rp + ib);
    double sv[BS];
    if (nk == 3) {
      double v[3][B2], x[3][BS], di[B2];
#pragma unroll
      for (int k = 0; k < 3; k++) ilu_ld_block<BS>(bp + k * B2, v[k]);
      if (bwd) ilu_ld_block<BS>(bp + 3 * B2, di);
      ilu_ld_vec<BS>(zb, i0.x, sv);
      ilu_ld_vec<BS>(zb, i0.y, x[0]);
      ilu_ld_vec<BS>(zb, i0.z, x[1]);
      ilu_ld_vec<BS>(zb, i0.w, x[2]);
      // three independent block products, then (s - p0) - (p1 + p2): the FP64 dependency chain of the level is
      // bs + 2 operations deep instead of 3 bs (the sweep is a chain of dependent levels run by a few warps, so its
      // time is this latency, not throughput).  Same products as the sequential MatSolve loop, the sum associated
      // differently: results agree to rounding.
      double pr[3][BS];
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int i = 0; i < BS; i++) {
          double acc = v[k][i] * x[k][0];
#pragma unroll
          for (int j = 1; j < BS; j++) acc += v[k][j * BS + i] * x[k][j];
          pr[k][i] = acc;
        }
#pragma unroll
      for (int i = 0; i < BS; i++) sv[i] = (sv[i] - pr[0][i]) - (pr[1][i] + pr[2][i]);
      if (bwd) {
        double t[BS];
#pragma unroll
        for (int i = 0; i < BS; i++) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < BS; j++) acc += di[j * BS + i] * sv[j];
          t[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < BS; i++) sv[i] = t[i];
      }
    } else {
      ilu_ld_vec<BS>(zb, i0.x, sv);
      const int *ip = reinterpret_cast<const int *>(rp);
      for (int k = 0; k < nk; k++) {
        double v[B2], xk[BS];
        ilu_ld_block<BS>(bp + k * B2, v);
        ilu_ld_vec<BS>(zb, ip[k + 1], xk);
#pragma unroll
        for (int j = 0; j < BS; j++)
#pragma unroll
          for (int i = 0; i < BS; i++) sv[i] -= v[j * BS + i] * xk[j];
      }
      if (bwd) {
        double di[B2], t[BS];
        ilu_ld_block<BS>(bp + nk * B2, di);
#pragma unroll
        for (int i = 0; i < BS; i++) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < BS; j++) acc += di[j * BS + i] * sv[j];
          t[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < BS; i++) sv[i] = t[i];
      }
    }
    if (BS == 2) {
      *reinterpret_cast<double2 *>(zb + i0.x) = make_double2(sv[0], sv[1]);
    } else {
#pragma unroll
      for (int i = 0; i < BS; i++) reinterpret_cast<double *>(zb + i0.x)[i] = sv[i];
    }
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Persistent-style grid for the streaming vector kernels: 4 CTAs of 256 threads per SM.
#define RED_BLOCKS (4 * WB_NUM_SMS)
#define KRY_MAXV 32  // vectors per fused multi-dot / multi-axpy launch (>= restart + 1 is not needed: one cycle
                     // of GMRES(30) dots against at most 30 vectors; longer restarts go in chunks)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;
}

struct KspState {
  double res, rnorm0;
  int its, reason, it_inner, pad;
};

// Arnoldi column `it_inner` is complete (hcol = dots, scal[0] = |w|^2): update the Givens QR and the
// convergence state (KSPGMRESUpdateHessenberg + KSPConvergedDefault).  One thread.
struct GmresUpd {
  double *hcol, *H, *cs, *sn, *rs, *scal;
  KspState *st;
  int *done;
  double rtol, atol, dtol;
  int m, maxit;
};
static __device__ void gmres_update(const GmresUpd &u) {
  if (*u.done) return;
  KspState *st = u.st;
  const int it = st->it_inner, m = u.m;
  const double tt = sqrt(u.scal[0]);
  u.hcol[it + 1] = tt;
  const bool happy = (tt < 1.e-30 * fmax(st->res, 1e-300)) || tt == 0.0;
  u.scal[1] = happy ? 1.0 : 1.0 / tt;
  double *Hc = u.H + (size_t)(m + 1) * it;
  for (int j = 0; j <= it + 1; j++) Hc[j] = u.hcol[j];
  for (int j = 0; j < it; j++) {
    const double t1 = Hc[j], t2 = Hc[j + 1];
    Hc[j] = u.cs[j] * t1 + u.sn[j] * t2;
    Hc[j + 1] = -u.sn[j] * t1 + u.cs[j] * t2;
  }
  const double hh = Hc[it], hp = Hc[it + 1];
  const double den = sqrt(hh * hh + hp * hp);
  if (den == 0.0) {
    st->reason = -5;  // KSP_DIVERGED_BREAKDOWN
    *u.done = 1;
    return;
  }
  u.cs[it] = hh / den;
  u.sn[it] = hp / den;
  u.rs[it + 1] = -u.sn[it] * u.rs[it];
  u.rs[it] = u.cs[it] * u.rs[it];
  Hc[it] = u.cs[it] * hh + u.sn[it] * hp;
  Hc[it + 1] = 0.0;
  const double res = fabs(u.rs[it + 1]);
  st->res = res;
  st->it_inner = it + 1;
  st->its += 1;
  int reason = 0;
  const double ttol = fmax(u.rtol * st->rnorm0, u.atol);
  if (res != res) reason = -9;
  else if (res <= ttol) reason = (res < u.atol) ? 3 : 2;
  else if (res >= u.dtol * st->rnorm0) reason = -4;
  if (!reason && happy) reason = 5;
  if (!reason && st->its >= u.maxit) reason = -3;
  if (reason) {
    st->reason = reason;
    *u.done = 1;
  }
}

#define WB_PROF_WORDS (16 * 256)
#define WB_LL_BYTES ((2 * KRY_MAXV * WB_NUM_SMS + 2 * KRY_MAXV) * 16)
struct KspWork {
  wb_ctx *ctx = nullptr;
  size_t n = 0, ld = 0;  // ld: n rounded up to 32 doubles so every basis vector is 256-byte aligned
  size_t cap = 0;        // allocated leading dimension: systems of different size (Jacobian, tracers) share the work space
  int m = 0;
  double *V = nullptr, *tmp = nullptr, *wbuf = nullptr, *small = nullptr, *part = nullptr;
  KspState *d_st = nullptr, *h_st = nullptr;
  int *d_done = nullptr;
  unsigned *d_counter = nullptr;
  int *d_bar = nullptr;                  // fused kernel: abort flag ([2])
  unsigned char *d_ll = nullptr;         // fused kernel: LL partials / totals of its grid-wide reductions
  unsigned long long *d_prof = nullptr;  // fused kernel: per-phase nanoseconds (+ iteration count) of every CTA, [cta][8]
};

// ---- sliced-ELL copy of a BAIJ matrix (slices of 32 rows, one warp each, thread per row) --------------------------
// Per slice one contiguous chunk: [nk][32] int32 column indices, then [nk][planes][32] value planes (double2 planes for
// even bs*bs): lane l of a warp reads element l of every plane -- fully coalesced from HBM, conflict-free from shared
// memory -- with arithmetic offsets: no row pointers, no dependent load before the column indices.  Rows shorter than
// the slice's widest are padded with zero blocks.  Used by the stand-alone SpMV (natural ordering) and, in the
// sub-domain-major ordering, by the persistent GMRES kernel (one bulk copy per slice).
#define WB_SELL_SLICE 32
// numeric part: BAIJ values -> plane layout of the slices (zero blocks where a row is shorter than its slice)
template <int BS>
__global__ void k_sell_fill(const double *__restrict__ val, const int4 *__restrict__ slices, int nslices,
                            const int32_t *__restrict__ ssrc, const int32_t *__restrict__ slot0,
                            unsigned char *__restrict__ sell) {
  constexpr int B2 = BS * BS, PW = IluPlane<BS>::PW, NPL = IluPlane<BS>::NP;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const int4 S = slices[s];
  const int nk = S.y >> 8;
  double *vals = reinterpret_cast<double *>(sell + (size_t)S.z * 16 + (size_t)nk * WB_SELL_SLICE * 4);
  const int32_t *src_ = ssrc + slot0[s];
  for (int k = 0; k < nk; k++) {
    const int src = src_[k * WB_SELL_SLICE + lane];
    double v[B2];
#pragma unroll
    for (int q = 0; q < B2; q++) v[q] = src >= 0 ? __ldcs(val + (size_t)src * B2 + q) : 0.0;
    double *dst = vals + (size_t)k * B2 * WB_SELL_SLICE;
#pragma unroll
    for (int q = 0; q < NPL; q++)
#pragma unroll
      for (int w = 0; w < PW; w++) dst[((size_t)q * WB_SELL_SLICE + lane) * PW + w] = v[q * PW + w];
  }
}


struct WbSell {
  int nslices = 0;
  int4 *d_slice = nullptr;  // (first row, rows | nk << 8, byte offset / 16, bytes)
  int32_t *d_ssrc = nullptr, *d_slot0 = nullptr;
  unsigned char *d_data = nullptr;
  uint64_t version = ~0ull;  // value version of the matrix the copy holds
};
// y = A (x * scale) through the sliced-ELL copy (built on first use, refreshed when the values changed); returns 1 if
// the matrix cannot use it (fall back to the BAIJ kernel)
int wb_sell_spmv(wb_mat *A, const double *d_x, const double *xg, const double *d_scale, double *d_xn, double *d_y,
                 const int *done);
void wb_sell_free(wb_mat *A);

// Krylov work space of a context (created on first use, shared by all its systems)
int wb_ensure_work(wb_ctx *c, size_t n, int m, KspWork **out);
KspWork *wb_find_work(wb_ctx *c);  // null if the context has not solved anything yet
int wb_fetch_state(KspWork &w);  // device solver state -> pinned host copy (synchronises); checks the P2P error flag

// ---- sub-domain-resident persistent GMRES (wb_fused.cu)
// symbolic part, after build_block_streams: CTA assignment, permuted sliced-ELL copy of the matrix, ring plans
int wb_fused_build(wb_pc *pc, const std::vector<int32_t> &blk_of);
int wb_fused_refresh(wb_pc *pc);  // numeric part: matrix values -> the permuted copy (after every Jacobian update)
void wb_fused_free(wb_pc *pc);
bool wb_fused_usable(const wb_mat *A, const wb_pc *pc, const wb_ksp_opts *o);
int wb_gmres_fused(wb_mat *A, wb_pc *pc, const wb_ksp_opts *o, const double *d_b, double *d_x, int *its, int *reason,
                   double *rnorm);
