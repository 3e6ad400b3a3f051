// wb_fused.cu -- sub-domain-resident persistent GMRES: KSPSolve_GMRES (restarted, classical Gram-Schmidt, left
// preconditioning with block-Jacobi / ILU(0)) as ONE cooperative kernel for the whole solve.
//
// Stands in for the same PETSc library code as wb_linalg.cu (KSPSolve_GMRES, MatMult_SeqBAIJ / MatMult_MPIBAIJ,
// MatSolve_SeqBAIJ_N_NaturalOrdering, VecMDot / VecMAXPY / VecNorm; call site src/timestepper.F90:1645-1836) with
// the same arithmetic per operation; what changes is where the data lives between the operations.
//
// One CTA per SM owns a fixed, contiguous set of block-Jacobi sub-domains.  The solver works in the sub-domain-major
// ordering (a symmetric permutation of the system: rows of a sub-domain contiguous, ascending natural order inside
// it -- the ordering the ILU(0) factors were computed in), so every vector segment a CTA touches is one contiguous
// span.  Per Krylov iteration a CTA runs, for each of its sub-domains,
//     SpMV rows (sliced-ELL copy of the BAIJ matrix, thread per row, operand gathered through L2)  ->  shared memory
//     ILU(0) forward / backward sweeps in shared memory (level records streamed by TMA bulk copies into a byte ring)
// then the Gram-Schmidt dots of its rows against the basis, a grid-wide reduction, the multi-AXPY + norm of its rows,
// and a second grid-wide reduction whose last CTA runs the Givens / convergence update.  The only grid-wide
// synchronisations of an iteration are those two reductions; kernel boundaries, the matrix-vector product's trip
// through HBM, the separate normalisation and the PC's read of its right-hand side are gone.
//
// Multi-GPU (one process per GPU, NVLink P2P over CUDA-IPC-mapped regions): boundary entries of the new Krylov vector
// are stored straight into the neighbours' ghost buffers by the CTA that owns them, UNSCALED, as soon as the
// multi-AXPY has produced them (the normalisation factor is a global scalar every rank applies itself), so the halo
// travels while the norm is being reduced; the two reductions are all-gathers of the per-GPU sums into every rank's
// slot, summed in rank order.  No collective launch, no host involvement until the solve is over.
#include <algorithm>
#include <math.h>
#include <stdlib.h>

#include "wb_linalg.cuh"

#define FZ_NBAR 32   // mbarriers per group and direction: at most FZ_NBAR - 1 level records in flight
#define FZ_MAXG 4    // sub-domain groups per CTA
#define FZ_SLICE WB_SELL_SLICE  // rows per sliced-ELL slice (one warp)
#define FZ_CH 8      // blocks of a row in flight per SpMV round
#define FZ_PYTH_ORTH 1e-6  // largest estimated relative error of that norm^2 from the basis' loss of orthogonality
#define FZ_PYTH_MIN 1e-4  // smallest |w'|^2 / |w|^2 for which the norm is taken from the dots (see k_gmres_fused)
#define FZ_SPIN_LIMIT 20000000000LL  // cycles (~10 s): a lost CTA / peer raises the abort flag instead of hanging the GPU

struct WbFusedPlan {
  int ncta = 0, ng = 1, gt = 512, nc = 512;  // CTAs, groups per CTA, threads per group, consumer threads
  int lt = 32, off_gm = 0, off_prof = 0;     // threads of a group in the level sweeps; offsets of the GMRES state, timers
  int sd_cap = 0, slice_cap = 0, rec_cap = 0;
  int zs_words = 0, ring_bytes = 0;
  size_t smem = 0;
  int off_red = 0, off_misc = 0, off_sd = 0, off_slice = 0, off_rec = 0, off_zs = 0, off_ring = 0;
  int nslots = 0, nslices = 0;  // sliced-ELL entries (slice, k, lane), slices
  int4 *d_cta = nullptr, *d_sd = nullptr, *d_slice = nullptr, *d_recA = nullptr, *d_recB = nullptr;
  int32_t *d_rec_ptr = nullptr, *d_ssrc = nullptr, *d_slot0 = nullptr;
  unsigned char *d_sell = nullptr;  // slices: [nk][32] column indices, then [nk][planes][32] values; one bulk copy each
  std::vector<int32_t> h_invperm;  // original row -> row in sub-domain-major order
  std::vector<int4> h_cta;
  // multi-GPU: boundary rows each CTA pushes (built on first use, when the halo plan and the peer map exist)
  bool push_built = false;
  int32_t *d_push_ptr = nullptr, *d_push_row = nullptr, *d_push_rank = nullptr, *d_push_off = nullptr;
};

// ================================================================ host: symbolic plan

template <class T> static int up(T **p, const std::vector<T> &v) {
  WB_CUDA(cudaMalloc((void **)p, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) WB_CUDA(wb_memcpy_sync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

void wb_fused_free(wb_pc *pc) {
  WbFusedPlan *f = pc->fused;
  if (!f) return;
  void *ptrs[] = {f->d_cta, f->d_sd, f->d_slice, f->d_recA, f->d_recB, f->d_rec_ptr, f->d_ssrc, f->d_slot0, f->d_sell,
                  f->d_push_ptr, f->d_push_row, f->d_push_rank, f->d_push_off};
  for (void *p : ptrs) cudaFree(p);
  delete f;
  pc->fused = nullptr;
}

static int g_fused_mode = -1;  // WB_FUSED = 0 turns the persistent kernel off (launch-per-operation GMRES of wb_linalg.cu)
static int fused_mode() {
  if (g_fused_mode < 0) {
    const char *e = getenv("WB_FUSED");
    g_fused_mode = e ? atoi(e) : 1;
  }
  return g_fused_mode;
}
// Norm of the new Krylov vector in the persistent kernel: 0 = always a second reduction (VecNorm, the reference's
// arithmetic); 1 (default) = from the dots (|w|^2 - sum h_j^2) when the solve spans several GPUs, where it replaces an
// NVLink all-gather by a barrier inside each GPU; 2 = always.  WB_FUSED_NORM overrides the default.
static int g_fused_norm = -1;
extern "C" int wb_ksp_set_fused_norm(int mode) {
  g_fused_norm = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
  return 0;
}

extern "C" int wb_ksp_set_fused(int on) {
  g_fused_mode = on < 0 ? 0 : on;  // 0 off, 1 automatic, 2 wherever it is usable
  return 0;
}

int wb_fused_build(wb_pc *pc, const std::vector<int32_t> &blk_of) {
  (void)blk_of;
  wb_mat *A = pc->A;
  wb_ctx *c = A->ctx;
  if (!fused_mode()) return 0;
  const int nb = pc->nb, bs = pc->bs, b2 = bs * bs, nsd = pc->nblk;
  if (nsd < 1 || nb < 1) return 0;
  cudaDeviceProp prop;
  WB_CUDA(cudaGetDeviceProperties(&prop, c->device));
  const int nsm = prop.multiProcessorCount;
  const size_t smem_max = prop.sharedMemPerBlockOptin;
  WbFusedPlan *f = new WbFusedPlan();
  f->ncta = std::min(std::min(nsd, nsm), WB_NUM_SMS);
  // contiguous runs of sub-domains per CTA, balanced by rows
  std::vector<int> first_sd(f->ncta + 1, 0);
  {
    int sd = 0;
    long long acc = 0;
    for (int ct = 0; ct < f->ncta; ct++) {
      first_sd[ct] = sd;
      const long long target = (long long)nb * (ct + 1) / f->ncta;
      const int left_ctas = f->ncta - ct - 1;
      // at least one sub-domain, and leave at least one for every remaining CTA
      do {
        acc += pc->h_blk[sd].y;
        sd++;
      } while (sd < nsd - left_ctas && acc + pc->h_blk[sd].y / 2 <= target);
    }
    first_sd[f->ncta] = nsd;
    if (sd != nsd) {  // the tail went to nobody: cannot happen with the loop above, but never launch a wrong plan
      delete f;
      return 0;
    }
  }
  int spc = 0;
  for (int ct = 0; ct < f->ncta; ct++) spc = std::max(spc, first_sd[ct + 1] - first_sd[ct]);
  f->sd_cap = spc;
  // 12 consumer warps + the producer warp = 13 warps: at most 4 per SM sub-partition, 128 registers per thread
  // (bs = 3: 6 + 1 warps, 2 per sub-partition, 255 registers: its 3x3 blocks need them)
  f->nc = bs >= 3 ? 192 : 384;
  // permutation
  f->h_invperm.assign(nb, 0);
  for (int p = 0; p < nb; p++) f->h_invperm[pc->h_blk_rows[p]] = p;
  // ---- sliced-ELL copy of the matrix in the sub-domain-major ordering
  std::vector<int4> sdtab(nsd), slices;
  std::vector<int32_t> ssrc, slot0;
  std::vector<unsigned char> sell;
  f->h_cta.assign(f->ncta, make_int4(0, 0, 0, 0));
  int max_slice_bytes = 0;
  for (int ct = 0; ct < f->ncta; ct++) {
    for (int sd = first_sd[ct]; sd < first_sd[ct + 1]; sd++) {
      const int row0 = pc->h_blk[sd].x, nr = pc->h_blk[sd].y;
      const int ns = (nr + FZ_SLICE - 1) / FZ_SLICE;
      sdtab[sd] = make_int4(row0, nr, (int)slices.size(), ns);
      for (int sidx_ = 0; sidx_ < ns; sidx_++) {
        const int r0 = row0 + sidx_ * FZ_SLICE, n = std::min(FZ_SLICE, row0 + nr - r0);
        int nk = 0;
        for (int l = 0; l < n; l++) {
          const int orow = pc->h_blk_rows[r0 + l];
          nk = std::max(nk, A->h_rowptr[orow + 1] - A->h_rowptr[orow]);
        }
        nk = std::max(nk, 1);
        const size_t bytes = (size_t)nk * (FZ_SLICE * 4 + (size_t)b2 * FZ_SLICE * 8);
        const size_t base = sell.size();
        if (base + bytes >= ((size_t)1 << 35) || nk > 255) {
          delete f;
          return 0;  // offsets are kept in 16-byte units in 32 bits
        }
        slices.push_back(make_int4(r0, n | (nk << 8), (int)(base / 16), (int)bytes));
        max_slice_bytes = std::max(max_slice_bytes, (int)bytes);
        sell.resize(base + bytes, 0);
        int32_t *ip = reinterpret_cast<int32_t *>(sell.data() + base);
        const size_t sbase = ssrc.size();
        slot0.push_back((int32_t)sbase);
        ssrc.resize(sbase + (size_t)nk * FZ_SLICE);
        for (int k = 0; k < nk; k++)
          for (int l = 0; l < FZ_SLICE; l++) {
            int col = r0 + std::min(l, n - 1), src = -1;  // padding: a zero block times the row's own entry
            if (l < n) {
              const int orow = pc->h_blk_rows[r0 + l];
              const int e = A->h_rowptr[orow] + k;
              if (e < A->h_rowptr[orow + 1]) {
                const int oc = A->h_colidx[e];
                col = oc < nb ? f->h_invperm[oc] : oc;  // ghost columns keep their index (>= nb)
                src = e;
              }
            }
            ip[(size_t)k * FZ_SLICE + l] = col;
            ssrc[sbase + (size_t)k * FZ_SLICE + l] = src;
          }
      }
    }
    f->h_cta[ct] = make_int4(first_sd[ct], first_sd[ct + 1] - first_sd[ct], 0, 0);
  }
  f->nslots = (int)ssrc.size();
  f->nslices = (int)slices.size();
  // ---- groups, ring plans.  Try 4, 2, 1 groups per CTA until the byte ring holds at least three of the largest
  // level records (deep enough to hide the HBM latency behind the level-to-level dependency chain)
  int max_rec_bytes = 0, max_rows = 0, max_level_rows = 1;
  for (const int4 &L : pc->h_lev) {
    max_rec_bytes = std::max(max_rec_bytes, L.y);
    max_level_rows = std::max(max_level_rows, L.z);
  }
  max_rec_bytes = std::max(max_rec_bytes, max_slice_bytes);
  for (int sd = 0; sd < nsd; sd++) max_rows = std::max(max_rows, pc->h_blk[sd].y);
  f->zs_words = ((max_rows + 1) * bs + 1) & ~1;
  bool ok = false;
  std::vector<int32_t> rec_ptr;
  std::vector<int4> recA, recB;
  for (int ng = std::min(bs >= 3 ? 2 : FZ_MAXG, spc >= 4 ? 4 : (spc >= 2 ? 2 : 1)); ng >= 1 && !ok; ng >>= 1) {
    f->ng = ng;
    f->gt = f->nc / ng;
    f->lt = std::min(f->gt, (max_level_rows + 31) / 32 * 32);
    // records per group and iteration
    int rec_cap = 0;
    for (int ct = 0; ct < f->ncta; ct++)
      for (int g = 0; g < ng; g++) {
        int cnt = 0;
        for (int sd = first_sd[ct] + g; sd < first_sd[ct + 1]; sd += ng) cnt += pc->h_blk[sd].w + sdtab[sd].w;
        rec_cap = std::max(rec_cap, cnt);
      }
    f->rec_cap = rec_cap;
    size_t off = 2 * FZ_MAXG * FZ_NBAR * sizeof(uint64_t);  // full + empty mbarriers
    f->off_red = (int)off;
    off += (size_t)std::max(f->nc / 32, WB_P2P_MAX_RANKS) * KRY_MAXV * sizeof(double);  // warp partials / rank sums
    f->off_misc = (int)off;
    off += 512;  // coefficients (KRY_MAXV doubles), flags, phase timers, solver state
    f->off_prof = (int)off;
    off += 16 * sizeof(unsigned long long);
    f->off_gm = (int)off;
    off += (size_t)(4 * KRY_MAXV + 8 + (KRY_MAXV + 1) * KRY_MAXV) * sizeof(double);  // column, rotations, rs, H
    f->off_sd = (int)off;
    off += (size_t)f->sd_cap * 2 * sizeof(int4);
    f->off_slice = (int)off;  // (unused: the slice descriptors travel in the record tables)
    f->off_rec = (int)off;
    off += (size_t)ng * (rec_cap + 1) * 2 * sizeof(int4);  // + 1: the level loop reads one descriptor ahead
    off = (off + 127) & ~(size_t)127;
    f->off_zs = (int)off;
    off += (size_t)ng * f->zs_words * sizeof(double);
    off = (off + 127) & ~(size_t)127;
    f->off_ring = (int)off;
    if (off + (size_t)ng * 2 * max_rec_bytes + 1024 > smem_max) continue;
    f->ring_bytes = (int)(((smem_max - 1024 - off) / ng) & ~(size_t)127);
    f->ring_bytes = std::min(f->ring_bytes, 1 << 20);
    f->smem = off + (size_t)ng * f->ring_bytes;
    // ring placement of every group's records (one period = one Krylov iteration; the period restarts at offset 0)
    rec_ptr.assign((size_t)f->ncta * ng + 1, 0);
    recA.clear();
    recB.clear();
    bool lag_ok = true;
    for (int ct = 0; ct < f->ncta; ct++)
      for (int g = 0; g < ng; g++) {
        const size_t r0 = recA.size();
        std::vector<int> roff, rbytes;
        int pos = 0;
        for (int sd = first_sd[ct] + g, slot = g; sd < first_sd[ct + 1]; sd += ng, slot += ng) {
          const int lev0 = pc->h_blk[sd].z, nl = pc->h_blk[sd].w;
          // the sub-domain's matrix slices (SpMV), then its level records (ILU sweeps), in the order they are consumed
          for (int sl = 0; sl < sdtab[sd].w; sl++) {
            const int4 S = slices[sdtab[sd].z + sl];
            const int bytes = (S.w + 15) & ~15;
            if (pos + bytes > f->ring_bytes) pos = 0;
            roff.push_back(pos);
            rbytes.push_back(bytes);
            recA.push_back(make_int4(S.z, S.w, pos, 0));
            recB.push_back(make_int4(S.x, S.y, slot, 4));
            pos += bytes;
          }
          for (int l = 0; l < nl; l++) {
            const int4 L = pc->h_lev[lev0 + l];
            const int bytes = (L.y + 15) & ~15;
            if (pos + bytes > f->ring_bytes) pos = 0;
            roff.push_back(pos);
            rbytes.push_back(bytes);
            recA.push_back(make_int4(L.x, L.y, pos, 0));
            recB.push_back(make_int4(L.z, L.w, slot, (l == 0 ? 1 : 0) | (l == nl - 1 ? 2 : 0)));
            pos += bytes;
          }
        }
        const int P = (int)roff.size();
        for (int i = 0; i < P; i++) {
          // the most recent earlier record (cyclically) whose space overlaps record i must have been consumed
          int lag = P;
          for (int d = 1; d < P; d++) {
            const int j = ((i - d) % P + P) % P;
            if (roff[j] < roff[i] + rbytes[i] && roff[i] < roff[j] + rbytes[j]) {
              lag = d;
              break;
            }
          }
          recA[r0 + i].w = std::max(1, std::min(lag, FZ_NBAR - 1));
          // the sweep threads ask for record i + 1 before they release record i, so no record may take the space of its
          // predecessor: a ring of three largest records guarantees that, a smaller one is checked record by record
          if (P >= 3 && lag < 2) lag_ok = false;
        }
        rec_ptr[(size_t)ct * ng + g + 1] = (int32_t)recA.size();
      }
    if (!lag_ok) continue;  // fewer groups per CTA = larger rings
    ok = true;
  }
  if (getenv("WB_FUSED_VERBOSE"))
    fprintf(stderr, "[wb_fused] %s: %d sub-domains (max %d rows) on %d CTAs (max %d each), %d group(s) x %d threads, level records <= %d bytes, "
                    "ring %d bytes per group, %d records per group and iteration, shared memory %zu of %zu bytes\n",
            ok ? "plan" : "NOT USABLE (level records do not fit the ring)", nsd, max_rows, f->ncta, spc, f->ng, f->gt, max_rec_bytes,
            f->ring_bytes, f->rec_cap, f->smem, smem_max);
  if (!ok) {
    delete f;
    return 0;
  }
  WB_TRY(up(&f->d_cta, f->h_cta));
  WB_TRY(up(&f->d_sd, sdtab));
  WB_TRY(up(&f->d_slice, slices));
  WB_TRY(up(&f->d_recA, recA));
  WB_TRY(up(&f->d_recB, recB));
  WB_TRY(up(&f->d_rec_ptr, rec_ptr));
  WB_TRY(up(&f->d_ssrc, ssrc));
  WB_TRY(up(&f->d_slot0, slot0));
  WB_CUDA(cudaMalloc(&f->d_sell, sell.size() + WB_PAD_BYTES));
  WB_CUDA(wb_memcpy_sync(f->d_sell, sell.data(), sell.size(), cudaMemcpyHostToDevice));
  pc->fused = f;
  return 0;
}

int wb_fused_refresh(wb_pc *pc) {
  WbFusedPlan *f = pc->fused;
  if (!f || f->nslices == 0) return 0;
  wb_mat *A = pc->A;
  wb_ctx *c = A->ctx;
  const int grid = wb_grid((size_t)f->nslices * 32, 256);
  switch (pc->bs) {
    case 1: k_sell_fill<1><<<grid, 256, 0, c->stream>>>(A->d_val, f->d_slice, f->nslices, f->d_ssrc, f->d_slot0, f->d_sell); break;
    case 2: k_sell_fill<2><<<grid, 256, 0, c->stream>>>(A->d_val, f->d_slice, f->nslices, f->d_ssrc, f->d_slot0, f->d_sell); break;
    default: k_sell_fill<3><<<grid, 256, 0, c->stream>>>(A->d_val, f->d_slice, f->nslices, f->d_ssrc, f->d_slot0, f->d_sell); break;
  }
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

// ================================================================ device: the persistent kernel

struct FusedArgs {
  // plan
  const int4 *cta, *sd, *blk, *slice, *recA, *recB;
  const int32_t *rec_ptr, *perm;
  const unsigned char *sell;
  const double *stream;
  int ng, gt, nc, lt, sd_cap, slice_cap, rec_cap, zs_words, ring_bytes;
  int off_red, off_misc, off_gm, off_prof, off_sd, off_slice, off_rec, off_zs, off_ring;
  int nb, ncta;
  size_t ld;
  // vectors: b / xout in the caller's ordering, everything else in the sub-domain-major ordering
  const double *b;
  double *xout, *V, *wa, *wb, *xp;
  unsigned char *llpart;  // LL partials and totals of the grid-wide reductions (zeroed before every solve)
  int *bar;  // [0] arrivals, [1] release generation, [2] abort
  GmresUpd upd;
  double *yv;
  // multi-GPU
  WbP2PDev P;
  int nneigh;
  const int32_t *nb_rank;
  unsigned long long ll_off[WB_P2P_MAX_RANKS], ll_stride[WB_P2P_MAX_RANKS];  // LL ghost buffers of every rank
  const int32_t *push_ptr, *push_row, *push_rank, *push_off;
  int *fseq;
  unsigned long long *prof;
  int pyth;    // norm of the new Krylov vector from |w|^2 - sum h_j^2 (one machine-wide reduction per iteration)
  int jitter;  // test aid (WB_FUSED_JITTER = 2^k ns): random per-thread delays at every phase boundary
};

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking phase test (try_wait may suspend the thread for a while: the producer polls several barriers)
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// consumer-side wait for a level record: bounded, so that a broken ring plan raises the abort flag instead of hanging
// (abort_flag = &bar[2]; the first record wait that times out leaves a note for the host's error message in bar[8..11]:
// which wait, CTA, thread, the record number it waited for)
__device__ __noinline__ void fz_note_timeout(int *abort_flag, int code, int q) {
  int *d = abort_flag + 6;
  if (atomicCAS(d, 0, code) == 0) {
    d[1] = blockIdx.x;
    d[2] = threadIdx.x;
    d[3] = q;
  }
  atomicExch(abort_flag, 1);
}
__device__ __forceinline__ void fz_mbar_wait(uint64_t *bar, uint32_t parity, int *abort_flag, int q = -1) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++n & 1023u) == 0) {
      if (clock64() - t0 > FZ_SPIN_LIMIT) fz_note_timeout(abort_flag, 1, q);
      if (__ldcg(abort_flag)) return;
    }
  }
}
__device__ __forceinline__ int fz_ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fz_st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long fz_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- LL transport of one double: a 16-byte store of (low half | seq << 32, high half | seq << 32); the reader
// polls until both words carry the sequence number it expects -- data and flag travel together, no fence
__device__ __forceinline__ void ll_store(void *slot, double v, int seq) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v), sq = (unsigned long long)(unsigned)seq << 32;
  const unsigned long long w0 = (b & 0xffffffffull) | sq, w1 = (b >> 32) | sq;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ double ll_load_wait(const void *slot, int seq, int *err) {
  unsigned long long w0, w1;
  const unsigned sq = (unsigned)seq;
  long long t0 = 0;
  unsigned n = 0;
  while (true) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
    if ((unsigned)(w0 >> 32) == sq && (unsigned)(w1 >> 32) == sq) break;
    if ((++n & 1023u) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > FZ_SPIN_LIMIT) {
        atomicExch(err, 1);
        break;
      }
    }
  }
  return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}

// the same inside one GPU (partials and totals of the grid-wide reductions live in this GPU's memory): L2 is the point
// of coherence, so GPU-scope relaxed accesses are enough (plain cache-global loads are weak: a polling loop may never
// observe the store)
__device__ __forceinline__ void ll_store_gpu(void *slot, double v, int seq) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v), sq = (unsigned long long)(unsigned)seq << 32;
  const unsigned long long w0 = (b & 0xffffffffull) | sq, w1 = (b >> 32) | sq;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ double ll_load_wait_gpu(const void *slot, int seq, int *err) {
  unsigned long long w0, w1;
  const unsigned sq = (unsigned)seq;
  long long t0 = 0;
  unsigned n = 0;
  while (true) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
    if ((unsigned)(w0 >> 32) == sq && (unsigned)(w1 >> 32) == sq) break;
    if ((++n & 1023u) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > FZ_SPIN_LIMIT) {
        atomicExch(err, 1);
        break;
      }
    }
  }
  return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}

#define FZ_BAR_ALL 15  // named barrier of all consumer threads; groups use 1 + g

// ---- grid-wide (and machine-wide) reductions without atomics or barriers --------------------------------------
// Every CTA contributes nv local sums; afterwards every CTA holds the nv global sums, bit-identical everywhere.
//   1. each CTA stores its partials in LL format (value and sequence number in one 16-byte word pair) into
//      part[kind][j][cta];
//   2. value j has an owner CTA (j mod #CTAs): it polls the partials of all CTAs, sums them in a fixed order and
//      publishes the total -- single GPU: into tot[kind][j]; multi-GPU: straight into every rank's NVLink slot;
//   3. every CTA polls the totals (multi-GPU: the slots of all ranks, summed in rank order).
// The flag travels with the data, so nothing waits for a fence or an atomic round trip: the latency is two
// store-to-poll hops inside the GPU (plus one NVLink flight).  `release` (used where the reduction also has to act as a
// grid barrier for the vector rows written before it) makes the CTA's earlier global writes visible first.
struct FzRed {
  unsigned char *part, *tot;  // [2][KRY_MAXV][WB_NUM_SMS] and [2][KRY_MAXV] LL words
  int ncta;
};
__device__ __forceinline__ void fz_reduce_ll(const FusedArgs &a, const FzRed &R, int kind, int seq, int nv, double *s_vals,
                                             double *s_tmp, bool release, bool multi, int mseq, size_t slot_off,
                                             int per_rank, int cta, int tid, int nc) {
  const int lane = tid & 31;
  if (release) __threadfence();    // every thread's earlier global writes (its rows of the vector) become visible first
  bar_sync_named(FZ_BAR_ALL, nc);  // s_vals complete
  if (tid < nv) ll_store_gpu(R.part + ((size_t)(kind * KRY_MAXV + tid) * WB_NUM_SMS + cta) * 16, s_vals[tid], seq);
  const size_t buf = multi ? (size_t)(mseq & 1) * WB_P2P_MAX_RANKS * per_rank * 16 : 0;
  for (int j = cta; j < nv; j += R.ncta) {  // owner duty
    if (tid < R.ncta)
      s_tmp[tid] = ll_load_wait_gpu(R.part + ((size_t)(kind * KRY_MAXV + j) * WB_NUM_SMS + tid) * 16, seq, &a.bar[2]);
    bar_sync_named(FZ_BAR_ALL, nc);
    if (tid < 32) {
      double t = 0.0;
      for (int c = lane; c < R.ncta; c += 32) t += s_tmp[c];  // fixed order: lane strides, then the shuffle tree
      t = warp_sum(t);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (!multi) {
        if (lane == 0) ll_store_gpu(R.tot + (size_t)(kind * KRY_MAXV + j) * 16, t, seq);
      } else if (lane < a.P.nranks) {
        ll_store(a.P.region[lane] + slot_off + buf + (size_t)(a.P.rank * per_rank + j) * 16, t, mseq);
      }
    }
    bar_sync_named(FZ_BAR_ALL, nc);
  }
  if (!multi) {
    if (tid < nv) s_vals[tid] = ll_load_wait_gpu(R.tot + (size_t)(kind * KRY_MAXV + tid) * 16, seq, &a.bar[2]);
  } else {
    const WbP2PDev &P = a.P;
    for (int idx = tid; idx < P.nranks * nv; idx += nc) {
      const int r = idx / nv, j = idx - r * nv;
      s_tmp[r * KRY_MAXV + j] = ll_load_wait(P.region[P.rank] + slot_off + buf + (size_t)(r * per_rank + j) * 16, mseq, P.err);
    }
    bar_sync_named(FZ_BAR_ALL, nc);
    if (tid < nv) {
      double t = 0.0;
      for (int r = 0; r < P.nranks; r++) t += s_tmp[r * KRY_MAXV + tid];
      s_vals[tid] = t;
    }
  }
  bar_sync_named(FZ_BAR_ALL, nc);
}

// The small GMRES state (Hessenberg matrix in its rotated form, Givens rotations, right-hand side of the least-squares
// problem, convergence state) lives in the shared memory of EVERY CTA and is updated redundantly by each of them from
// the same reduced numbers with the same instructions: no broadcast, no serial section behind a grid barrier.
struct FzState {
  double res, rnorm0, scal1;
  int its, it_inner, reason, done;
};
// Arnoldi column `it` complete: s_h[0..it] = dots, nrm2 = |w|^2.  One thread; the arithmetic of gmres_update
// (wb_linalg.cuh: KSPGMRESUpdateHessenberg + KSPConvergedDefault) on the CTA's copy of the state.
__device__ __forceinline__ void fz_gmres_update(const GmresUpd &u, FzState *st, double nrm2, double *s_h, double *s_H,
                                                double *s_cs, double *s_sn, double *s_rs) {
  const int it = st->it_inner, m = u.m;
  const double tt = sqrt(nrm2);
  s_h[it + 1] = tt;
  const bool happy = (tt < 1.e-30 * fmax(st->res, 1e-300)) || tt == 0.0;
  st->scal1 = happy ? 1.0 : 1.0 / tt;
  double *Hc = s_H + (size_t)(m + 1) * it;
  for (int j = 0; j <= it + 1; j++) Hc[j] = s_h[j];
  for (int j = 0; j < it; j++) {
    const double t1 = Hc[j], t2 = Hc[j + 1];
    Hc[j] = s_cs[j] * t1 + s_sn[j] * t2;
    Hc[j + 1] = -s_sn[j] * t1 + s_cs[j] * t2;
  }
  const double hh = Hc[it], hp = Hc[it + 1];
  const double den = sqrt(hh * hh + hp * hp);
  if (den == 0.0) {
    st->reason = -5;  // KSP_DIVERGED_BREAKDOWN
    st->done = 1;
    return;
  }
  s_cs[it] = hh / den;
  s_sn[it] = hp / den;
  s_rs[it + 1] = -s_sn[it] * s_rs[it];
  s_rs[it] = s_cs[it] * s_rs[it];
  Hc[it] = s_cs[it] * hh + s_sn[it] * hp;
  Hc[it + 1] = 0.0;
  const double res = fabs(s_rs[it + 1]);
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
    st->done = 1;
  }
}

// BCGS = 1: the same resident machinery running KSPSolve_BCGS instead of KSPSolve_GMRES (main loop at the end)
template <int BS, int BCGS = 0>
__global__ void __launch_bounds__(BS >= 3 ? 224 : 416, 1) k_gmres_fused(const FusedArgs a) {
  constexpr int B2 = BS * BS, PW = IluPlane<BS>::PW, NPL = IluPlane<BS>::NP;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, nc = a.nc, ng = a.ng, gt = a.gt;
  const int cta = blockIdx.x;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  double *s_red = reinterpret_cast<double *>(smem + a.off_red);       // [warps][KRY_MAXV]
  double *s_cf = reinterpret_cast<double *>(smem + a.off_misc);       // [KRY_MAXV + 2]
  double *s_h = reinterpret_cast<double *>(smem + a.off_gm);          // [KRY_MAXV + 2] dots / Hessenberg column
  double *s_cs = s_h + KRY_MAXV + 2, *s_sn = s_cs + KRY_MAXV, *s_rs = s_sn + KRY_MAXV;  // rs: [KRY_MAXV + 1]
  double *s_H = s_rs + KRY_MAXV + 2;                                  // [(m + 1) * m]
  double *s_pn = reinterpret_cast<double *>(smem + a.off_misc + 280);  // [3] norm^2 from the dots, "use it" flag, cycle-start residual
  FzState *s_st = reinterpret_cast<FzState *>(smem + a.off_misc + 448);
  volatile int *s_stop = reinterpret_cast<volatile int *>(smem + a.off_misc + 324);
  int *s_nrec = reinterpret_cast<int *>(smem + a.off_misc + 328);     // [FZ_MAXG]
  volatile int *s_consumed = reinterpret_cast<volatile int *>(smem + a.off_misc + 432);  // [FZ_MAXG] level records consumed
  int4 *s_sd = reinterpret_cast<int4 *>(smem + a.off_sd);             // [sd_cap][2]
  int4 *s_rec = reinterpret_cast<int4 *>(smem + a.off_rec);           // [ng][rec_cap][2]
  const int4 C = a.cta[cta];
  const int sd0 = C.x, nsd = C.y;

  // ---- set-up: barriers, tables -> shared memory
  if (tid == 0) {
    for (int i = 0; i < FZ_MAXG * FZ_NBAR; i++) {
      mbar_init(&full[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *s_stop = 0;
    for (int i = 0; i < FZ_MAXG; i++) s_consumed[i] = 0;
  }
  for (int s = tid; s < nsd; s += blockDim.x) {
    s_sd[2 * s] = a.sd[sd0 + s];
    s_sd[2 * s + 1] = a.blk[sd0 + s];
  }
  for (int g = 0; g < ng; g++) {
    const int r0 = a.rec_ptr[cta * ng + g], r1 = a.rec_ptr[cta * ng + g + 1];
    if (tid == 0) s_nrec[g] = r1 - r0;
    for (int r = tid; r < r1 - r0; r += blockDim.x) {
      s_rec[((size_t)g * (a.rec_cap + 1) + r) * 2] = a.recA[r0 + r];
      s_rec[((size_t)g * (a.rec_cap + 1) + r) * 2 + 1] = a.recB[r0 + r];
    }
  }
  __syncthreads();

  // ================================================================ producer warp: feeds the rings of all groups.
  // Every LANE owns every (32 / ng)-th record of one group: it waits until the record whose ring space its next record
  // takes has been consumed, issues the bulk copy, and moves on -- 32 records are being issued at any time, so the
  // per-record issue cost (a few hundred cycles of dependent instructions in one thread) is off the consumers'
  // critical path.  The lanes run ahead of the consumers across phases and iterations until the stop flag is raised.
  if (tid >= nc) {
    const int pl = tid - nc, g = pl % ng, per = 32 / ng;
    const int P = s_nrec[g];
    if (P > 0) {
      const int4 *rec = s_rec + (size_t)g * (a.rec_cap + 1) * 2;
      uint64_t *fullg = full + g * FZ_NBAR;
      unsigned char *ring = smem + a.off_ring + (size_t)g * a.ring_bytes;
      long long q = pl / ng, qlast = -1;
      int pos = (int)(q % P);
      while (true) {
        const int4 A = rec[2 * pos];
        const bool is_slice = (rec[2 * pos + 1].w & 4) != 0;
        bool stopped = false;
        if (q >= A.w) {
          // the record whose space this one takes must have been consumed.  The consumers publish a monotonic count
          // (a lane may be asking about a record far ahead of the consumers: mbarrier phase parities would alias)
          const int need = (int)(q - A.w) + 1;
          while (s_consumed[g] < need) {
            if (*s_stop) {
              stopped = true;
              break;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // their reads before this bulk write
        }
        if (stopped || *s_stop) break;
        uint64_t *fb = &fullg[q % FZ_NBAR];
        mbar_expect_tx(fb, (uint32_t)A.y);
        tma_load_1d(ring + A.z, is_slice ? static_cast<const void *>(a.sell + (size_t)A.x * 16)
                                         : static_cast<const void *>(a.stream + A.x), (uint32_t)A.y, fb);
        qlast = q;
        q += per;
        pos = (int)((pos + per) % P);
      }
      // the copy this lane issued last must have landed before the CTA's shared memory goes away (its earlier ones
      // were consumed: the lane waited for a later record's space)
      if (qlast >= 0) mbar_wait(&fullg[qlast % FZ_NBAR], (uint32_t)((qlast / FZ_NBAR) & 1));
    }
    return;
  }

  // ================================================================ consumers
  const int g = tid / gt, gtid = tid - g * gt;
  const int gwarp = gtid >> 5, gwarps = gt >> 5, lane = tid & 31;
  const int warp = tid >> 5, nwarps = nc >> 5;
  double *zs = reinterpret_cast<double *>(smem + a.off_zs) + (size_t)g * a.zs_words;
  const unsigned char *ring = smem + a.off_ring + (size_t)g * a.ring_bytes;
  const int4 *rec = s_rec + (size_t)g * (a.rec_cap + 1) * 2;
  uint64_t *fullg = full + g * FZ_NBAR;
  const int nrec = s_nrec[g];
  // threads of the group that take part in the level sweeps: the widest level rounded to warps, plus (when the group
  // has a warp to spare) one warp whose first lane only looks out for the next record
  const int lt = a.lt, ltw = lt + 32 <= gt ? lt + 32 : lt;
  const bool waiter = gtid == ltw - 32;
  // rows [R0, R1) and elements [e0, e1) of this CTA; vector body [eb, ee) is 16-byte aligned
  const int R0 = s_sd[0].x, R1 = s_sd[2 * (nsd - 1)].x + s_sd[2 * (nsd - 1)].y;
  const int e0 = R0 * BS, e1 = R1 * BS;
  const int eb = (e0 + 1) & ~1, ee = e1 & ~1;
  const bool multi = a.P.on != 0 && a.P.nranks > 1;
  const FzRed R = {a.llpart, a.llpart + (size_t)2 * KRY_MAXV * WB_NUM_SMS * 16, a.ncta};
  int rseqA = 0, rseqB = 0;  // sequence numbers of the dots (kind 0) and norm / barrier (kind 1) reductions
  int hseq = 0, aseq = 0, bseq = 0;
  if (multi) {
    hseq = __ldcg(&a.fseq[0]);
    aseq = __ldcg(&a.fseq[1]);
    bseq = __ldcg(&a.fseq[2]);
  }
  const GmresUpd &u = a.upd;
  const int m = u.m;
  int qrec = 0, ri = 0;  // records this group has consumed (slot = qrec % FZ_NBAR, parity = (qrec / FZ_NBAR) & 1), position in its list
  // phase timers of CTA 0 (thread 0), kept in shared memory
  unsigned long long *s_prof = reinterpret_cast<unsigned long long *>(smem + a.off_prof);  // [16]; [8] = previous stamp
  const bool pyth = a.pyth != 0;
  const bool profiler = (tid == 0);  // every CTA keeps its own phase times (CTA 0's are the ones the ABI reports)
  if (profiler) {
    for (int k = 0; k < 16; k++) s_prof[k] = 0;
    s_prof[8] = fz_now();
  }
#define FZ_STAMP(k)                  \
  do {                               \
    if (a.jitter) {                  \
      const unsigned h_ = ((unsigned)clock() + tid * 40503u + cta * 9973u) * 2654435761u; \
      __nanosleep((h_ >> 12) & (unsigned)(a.jitter - 1)); \
    }                                \
    if (profiler) {                  \
      const unsigned long long t_ = fz_now(); \
      s_prof[k] += t_ - s_prof[8];   \
      s_prof[8] = t_;                \
    }                                \
  } while (0)

  // x = 0 on the rows of this CTA
  for (int e = e0 + tid; e < e1; e += nc) a.xp[e] = 0.0;

  // ---- SpMV + ILU(0) of the CTA's sub-domains.  mode 0: t = b; 1: t = A (xop * s), V_it = xop * s;
  // 2: t = b - A xop.  Result w_dst = M^-1 t on the CTA's rows.
  auto sp_phase = [&](int mode, const double *xop, double s, double *vstore, double *w_dst) {
    // ghost entries: this GPU's LL buffer (hseq & 1); every entry carries the sequence number of its push, so a
    // gather simply polls the entry it needs -- rows without ghost columns never wait
    const unsigned char *xg = multi ? a.P.region[a.P.rank] + a.ll_off[a.P.rank] + (size_t)(hseq & 1) * a.ll_stride[a.P.rank]
                                    : nullptr;
    for (int sl = g; sl < nsd; sl += ng) {
      const int4 sA = s_sd[2 * sl], sB = s_sd[2 * sl + 1];
      const int row0 = sA.x, nr = sA.y, ns = sA.w, nl = sB.w;
      if (gtid < BS) zs[nr * BS + gtid] = 0.0;  // the slot padding blocks of the level records multiply
      // ---- SpMV: the sub-domain's matrix slices arrive through the ring (records qrec .. qrec + ns - 1, one warp
      // each, `gwarps` of them per round); the operand is gathered through L2
      for (int s0 = 0; s0 < ns; s0 += gwarps) {
        const int si = s0 + gwarp;
        if (si < ns) {
          const int qi = qrec + si;
          const int4 A = rec[2 * (ri + si)], B = rec[2 * (ri + si) + 1];
          fz_mbar_wait(&fullg[qi & (FZ_NBAR - 1)], (uint32_t)((qi / FZ_NBAR) & 1), &a.bar[2], qi);
          const int n = B.y & 255, nk = B.y >> 8, srow0 = B.x;
          const int row = srow0 + lane, li = row - row0;
          if (mode == 0) {
            if (lane < n) {
              const size_t ob = (size_t)a.perm[row] * BS;
#pragma unroll
              for (int i = 0; i < BS; i++) zs[li * BS + i] = a.b[ob + i];
            }
          } else {
            const int32_t *ip = reinterpret_cast<const int32_t *>(ring + A.z) + lane;
            const double *vp = reinterpret_cast<const double *>(ring + A.z + (size_t)nk * FZ_SLICE * 4) + (size_t)lane * PW;
            double acc[BS];
#pragma unroll
            for (int i = 0; i < BS; i++) acc[i] = 0.0;
            // FZ_CH blocks of the row at a time: their operand entries are gathered together (one L2 round trip)
            for (int k0 = 0; k0 < nk; k0 += FZ_CH) {
              int col[FZ_CH];
              double x[FZ_CH][BS];
#pragma unroll
              for (int uu = 0; uu < FZ_CH; uu++) col[uu] = ip[min(k0 + uu, nk - 1) * FZ_SLICE];
#pragma unroll
              for (int uu = 0; uu < FZ_CH; uu++) {
                const bool on = k0 + uu < nk;
                const bool own = col[uu] < a.nb;
                if (on && !own) {
                  const unsigned char *gp = xg + (size_t)(col[uu] - a.nb) * BS * 16;
#pragma unroll
                  for (int j = 0; j < BS; j++) x[uu][j] = ll_load_wait(gp + j * 16, hseq, a.P.err) * s;
                } else {
                  const double *xp_ = xop + (size_t)col[uu] * BS;
                  if (BS == 2) {
                    const double2 t = on ? __ldcg(reinterpret_cast<const double2 *>(xp_)) : make_double2(0.0, 0.0);
                    x[uu][0] = t.x * s;
                    x[uu][1] = t.y * s;
                  } else {
#pragma unroll
                    for (int j = 0; j < BS; j++) x[uu][j] = on ? __ldcg(xp_ + j) * s : 0.0;
                  }
                }
              }
#pragma unroll
              for (int uu = 0; uu < FZ_CH; uu++) {
                const int k = min(k0 + uu, nk - 1);  // past the end: the last block again, its x is zero
                const double *bp = vp + (size_t)k * B2 * FZ_SLICE;
                double v[B2];
#pragma unroll
                for (int qq = 0; qq < NPL; qq++) {
                  if (PW == 2) {
                    const double2 t = *reinterpret_cast<const double2 *>(bp + (size_t)qq * FZ_SLICE * 2);
                    v[2 * qq] = t.x;
                    v[2 * qq + 1] = t.y;
                  } else {
                    v[qq] = bp[(size_t)qq * FZ_SLICE];
                  }
                }
#pragma unroll
                for (int j = 0; j < BS; j++)
#pragma unroll
                  for (int i = 0; i < BS; i++) acc[i] += v[j * BS + i] * x[uu][j];
              }
            }
            if (lane < n) {
              if (mode == 1) {
#pragma unroll
                for (int i = 0; i < BS; i++) {
                  if (!BCGS) vstore[(size_t)row * BS + i] = __ldcg(xop + (size_t)row * BS + i) * s;
                  zs[li * BS + i] = acc[i];
                }
              } else {
                const size_t ob = (size_t)a.perm[row] * BS;
#pragma unroll
                for (int i = 0; i < BS; i++) zs[li * BS + i] = a.b[ob + i] - acc[i];
              }
            }
          }
        }
        bar_sync_named(1 + g, gt);
        if (gtid == 0) s_consumed[g] = qrec + min(s0 + gwarps, ns);  // the ring space of these slices may be reused
      }
      qrec += ns;
      ri += ns;
      if (profiler) s_prof[7] += fz_now() - s_prof[8];  // SpMV share of the phase (group 0)
      // ---- forward and backward sweeps, level by level, by the first `lt` threads of the group only (a level never
      // has more rows; the other warps would just burn issue slots walking the loop).  One thread looks out for the
      // next record while the others apply the current one; the level barrier then publishes its acquire to everybody.
      if (gtid < ltw && nl > 0) {
        int qi = qrec, rl = ri;
        fz_mbar_wait(&fullg[qi & (FZ_NBAR - 1)], (uint32_t)((qi / FZ_NBAR) & 1), &a.bar[2], qi);
        int4 A = rec[2 * rl], B = rec[2 * rl + 1];
        for (int l = 0; l < nl; l++) {
          const int4 An = rec[2 * (rl + 1)], Bn = rec[2 * (rl + 1) + 1];  // next record's descriptor (table is padded)
          if (gtid < lt)
            ilu_level<BS>(reinterpret_cast<const double *>(ring + A.z), B.x, B.y & 0xffff, (B.y >> 16) != 0, zs, lt, gtid);
          if (waiter && l + 1 < nl)
            fz_mbar_wait(&fullg[(qi + 1) & (FZ_NBAR - 1)], (uint32_t)(((qi + 1) / FZ_NBAR) & 1), &a.bar[2], qi + 1);
          bar_sync_named(8 + g, ltw);
          qi++;
          rl++;
          A = An;
          B = Bn;
          if (gtid == 0) s_consumed[g] = qi;  // the ring space of this record may be reused
        }
      }
      qrec += nl;
      ri += nl;
      if (ri >= nrec) ri = 0;  // a group's record list covers whole sub-domains: it ends exactly here
      bar_sync_named(1 + g, gt);
      for (int li = gtid; li < nr; li += gt) {
#pragma unroll
        for (int i = 0; i < BS; i++) w_dst[(size_t)(row0 + li) * BS + i] = zs[li * BS + i];
      }
      bar_sync_named(1 + g, gt);
    }
  };

  // ---- dots of w (this CTA's rows) against nv vectors V_j = V + j * ldv -> part[cta][j]
  auto dots = [&](const double *w, const double *V, size_t ldv, int nv, double *s_out, bool with_ww) {
    double accw = 0.0;  // with_ww: w . w on the same pass, stored as sum number nv
    for (int j0 = 0; j0 < nv; j0 += 8) {
      const int nvc = min(8, nv - j0);
      double acc[8];
#pragma unroll
      for (int k = 0; k < 8; k++) acc[k] = 0.0;
      for (int i = (eb >> 1) + tid; i < (ee >> 1); i += 2 * nc) {
        // two entries per thread in flight: 2 x 8 vector loads before the first use
        const int i2 = i + nc < (ee >> 1) ? i + nc : i;
        const double2 wi = __ldcg(reinterpret_cast<const double2 *>(w + 2 * (size_t)i));
        double2 wj = __ldcg(reinterpret_cast<const double2 *>(w + 2 * (size_t)i2));
        if (i2 == i) wj = make_double2(0.0, 0.0);
        double2 v[8], v2[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int j = j0 + (k < nvc ? k : nvc - 1);  // past nvc: re-read the last vector, sums dropped
          v[k] = *reinterpret_cast<const double2 *>(V + (size_t)j * ldv + 2 * (size_t)i);
          v2[k] = *reinterpret_cast<const double2 *>(V + (size_t)j * ldv + 2 * (size_t)i2);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
          acc[k] += wi.x * v[k].x;
          acc[k] += wi.y * v[k].y;
          acc[k] += wj.x * v2[k].x;
          acc[k] += wj.y * v2[k].y;
        }
        if (with_ww && j0 == 0) {
          accw += wi.x * wi.x;
          accw += wi.y * wi.y;
          accw += wj.x * wj.x;
          accw += wj.y * wj.y;
        }
      }
      if (tid == 0) {  // unaligned head / tail element
        if (eb > e0) {
          for (int k = 0; k < nvc; k++) acc[k] += w[e0] * V[(size_t)(j0 + k) * ldv + e0];
          if (with_ww && j0 == 0) accw += w[e0] * w[e0];
        }
        if (ee < e1 && ee >= eb) {
          for (int k = 0; k < nvc; k++) acc[k] += w[e1 - 1] * V[(size_t)(j0 + k) * ldv + e1 - 1];
          if (with_ww && j0 == 0) accw += w[e1 - 1] * w[e1 - 1];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const double sres = warp_sum(acc[k]);
        if (lane == 0 && k < nvc) s_red[warp * KRY_MAXV + j0 + k] = sres;
      }
    }
    if (with_ww) {
      const double sres = warp_sum(accw);
      if (lane == 0) s_red[warp * KRY_MAXV + nv] = sres;
    }
    bar_sync_named(FZ_BAR_ALL, nc);
    if (tid < nv + (with_ww ? 1 : 0)) {
      double sres = 0.0;
      for (int wq = 0; wq < nwarps; wq++) sres += s_red[wq * KRY_MAXV + tid];
      s_out[tid] = sres;
    }
  };

  // ---- w += sum_j s_cf[j] V_j on this CTA's rows (sequential in j per entry, as VecMAXPY); with `norm` the CTA's
  // part of |w|^2 -> part[cta][0]
  auto maxpy = [&](double *w, const double *V, size_t ldv, int nv, double *s_out) {
    double nrm = 0.0;
    for (int i = (eb >> 1) + tid; i < (ee >> 1); i += 2 * nc) {
      const int i2 = i + nc;
      const bool two = i2 < (ee >> 1);
      const int ib = two ? i2 : i;
      double2 wi = __ldcg(reinterpret_cast<const double2 *>(w + 2 * (size_t)i));
      double2 wj = __ldcg(reinterpret_cast<const double2 *>(w + 2 * (size_t)ib));
      for (int j0 = 0; j0 < nv; j0 += 8) {
        const int nvc = min(8, nv - j0);
        double2 v[8], v2[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int j = j0 + (k < nvc ? k : nvc - 1);
          v[k] = *reinterpret_cast<const double2 *>(V + (size_t)j * ldv + 2 * (size_t)i);
          v2[k] = *reinterpret_cast<const double2 *>(V + (size_t)j * ldv + 2 * (size_t)ib);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const double cj = k < nvc ? s_cf[j0 + k] : 0.0;
          wi.x += cj * v[k].x;
          wi.y += cj * v[k].y;
          wj.x += cj * v2[k].x;
          wj.y += cj * v2[k].y;
        }
      }
      *reinterpret_cast<double2 *>(w + 2 * (size_t)i) = wi;
      nrm += wi.x * wi.x;
      nrm += wi.y * wi.y;
      if (two) {
        *reinterpret_cast<double2 *>(w + 2 * (size_t)i2) = wj;
        nrm += wj.x * wj.x;
        nrm += wj.y * wj.y;
      }
    }
    if (tid == 0) {
      if (eb > e0) {
        double wi = w[e0];
        for (int j = 0; j < nv; j++) wi += s_cf[j] * V[(size_t)j * ldv + e0];
        w[e0] = wi;
        nrm += wi * wi;
      }
      if (ee < e1 && ee >= eb) {
        double wi = w[e1 - 1];
        for (int j = 0; j < nv; j++) wi += s_cf[j] * V[(size_t)j * ldv + e1 - 1];
        w[e1 - 1] = wi;
        nrm += wi * wi;
      }
    }
    if (s_out) {
      const double sres = warp_sum(nrm);
      if (lane == 0) s_red[warp * KRY_MAXV] = sres;
      bar_sync_named(FZ_BAR_ALL, nc);
      if (tid == 0) {
        double t = 0.0;
        for (int wq = 0; wq < nwarps; wq++) t += s_red[wq * KRY_MAXV];
        s_out[0] = t;
      }
    }
  };

  // ---- multi-GPU: the boundary rows of this CTA -> the neighbours' ghost buffers (buffer (hseq + 1) & 1), unscaled
  auto push = [&](const double *w) {
    if (!multi || a.nneigh == 0) return;
    bar_sync_named(FZ_BAR_ALL, nc);  // the rows were written by other threads of this CTA
    const int p0 = a.push_ptr[cta], p1 = a.push_ptr[cta + 1];
    for (int e = p0 + tid; e < p1; e += nc) {
      const int row = a.push_row[e], r = a.push_rank[e];
      unsigned char *ghost = a.P.region[r] + a.ll_off[r] + (size_t)((hseq + 1) & 1) * a.ll_stride[r] +
                             (size_t)a.push_off[e] * BS * 16;
#pragma unroll
      for (int i = 0; i < BS; i++) ll_store(ghost + i * 16, w[(size_t)row * BS + i], hseq + 1);
    }
  };
  if (BCGS) {
    // ================================================================ KSPSolve_BCGS (left preconditioning, x0 = 0,
    // convergence on the preconditioned residual norm): R, RP, P, V, S, T live in the first six columns of the basis
    // storage, every CTA updates its own rows; the scalars are replicated like the GMRES state.  Per iteration: two
    // products, three reductions of numbers ((V,RP); (T,S),(T,T); (R,RP),(R,R)) and two barriers that only order the
    // new operand rows inside the GPU (the neighbours' rows on other GPUs arrive as flagged halo entries).
    double *vR = a.V, *vRP = a.V + a.ld, *vP = a.V + 2 * a.ld, *vV = a.V + 3 * a.ld, *vS = a.V + 4 * a.ld,
           *vT = a.V + 5 * a.ld;
    double *sc = s_cs;  // [0] rho_old, [1] alpha, [2] omega_old, [3] beta, [4] omega, [5] rho, [6] flag: T . T == 0
    auto reduce_numbers = [&](int nv, bool release) {
      fz_reduce_ll(a, R, 0, ++rseqA, nv, s_h, s_red, release, multi, aseq + 1, WB_P2P_SLOT_FA, WB_P2P_MAXV, cta, tid, nc);
      aseq++;
    };
    auto barrier_gpu = [&]() {  // rows written by this GPU's CTAs become visible to each other
      if (tid == 0) s_cf[0] = 0.0;
      fz_reduce_ll(a, R, 1, ++rseqB, 1, s_cf, s_red, true, false, 0, 0, 1, cta, tid, nc);
    };
    if (tid == 0) {
      s_st->res = 0.0; s_st->rnorm0 = 0.0; s_st->scal1 = 1.0;
      s_st->its = 0; s_st->it_inner = 0; s_st->reason = 0; s_st->done = 0;
      sc[0] = 1.0; sc[1] = 1.0; sc[2] = 1.0; sc[6] = 0.0;
    }
    FZ_STAMP(5);
    sp_phase(0, a.xp, 1.0, nullptr, vR);  // R = M^-1 b
    FZ_STAMP(0);
    bar_sync_named(FZ_BAR_ALL, nc);
    for (int e = e0 + tid; e < e1; e += nc) {
      vRP[e] = __ldcg(vR + e);
      vP[e] = 0.0;
      vV[e] = 0.0;
    }
    dots(vR, vR, 0, 1, s_h, false);  // (R, R) = (R, RP)
    FZ_STAMP(1);
    reduce_numbers(1, false);
    if (tid == 0) {
      const double dp = sqrt(s_h[0]);
      s_h[1] = s_h[0];
      s_st->res = dp;
      s_st->rnorm0 = dp;
      int reason = 0;
      const double ttol = fmax(u.rtol * dp, u.atol);
      if (dp != dp) reason = -9;
      else if (dp <= ttol) reason = (dp < u.atol) ? 3 : 2;
      if (!reason && dp == 0.0) reason = 3;
      if (reason) {
        s_st->reason = reason;
        s_st->done = 1;
      }
    }
    bar_sync_named(FZ_BAR_ALL, nc);
    FZ_STAMP(2);
    while (!__ldcg(&a.bar[2]) && !s_st->done) {
      // ---- rho = (R, RP) (in s_h[0] from the last reduction), beta, P = R + beta (P - omega_old V)
      if (tid == 0) {
        const double rho = s_h[0];
        sc[5] = rho;
        if (rho == 0.0) {
          s_st->reason = -5;  // KSP_DIVERGED_BREAKDOWN
          s_st->done = 1;
        }
        sc[3] = (rho / sc[0]) * (sc[1] / sc[2]);
      }
      bar_sync_named(FZ_BAR_ALL, nc);
      if (s_st->done) break;
      {
        const double beta = sc[3], omegaold = sc[2];
        auto fp = [&](double r, double p_, double v) { return __dadd_rn(r, __dmul_rn(beta, __dsub_rn(p_, __dmul_rn(omegaold, v)))); };
        for (int i = (eb >> 1) + tid; i < (ee >> 1); i += nc) {  // 16-byte body of the CTA's rows, two entries per thread
          const double2 r = __ldcg(reinterpret_cast<const double2 *>(vR) + i), p_ = __ldcg(reinterpret_cast<const double2 *>(vP) + i),
                        v = __ldcg(reinterpret_cast<const double2 *>(vV) + i);
          reinterpret_cast<double2 *>(vP)[i] = make_double2(fp(r.x, p_.x, v.x), fp(r.y, p_.y, v.y));
        }
        if (tid == 0) {  // unaligned head / tail element
          if (eb > e0) vP[e0] = fp(__ldcg(vR + e0), __ldcg(vP + e0), __ldcg(vV + e0));
          if (ee < e1 && ee >= eb) vP[e1 - 1] = fp(__ldcg(vR + e1 - 1), __ldcg(vP + e1 - 1), __ldcg(vV + e1 - 1));
        }
      }
      push(vP);
      FZ_STAMP(3);
      barrier_gpu();
      hseq++;
      FZ_STAMP(4);
      if (__ldcg(&a.bar[2])) break;
      // ---- V = M^-1 A P, alpha = rho / (V, RP)
      sp_phase(1, vP, 1.0, nullptr, vV);
      FZ_STAMP(0);
      bar_sync_named(FZ_BAR_ALL, nc);
      dots(vV, vRP, 0, 1, s_h, false);
      FZ_STAMP(1);
      reduce_numbers(1, false);
      FZ_STAMP(2);
      if (__ldcg(&a.bar[2])) break;
      if (tid == 0) {
        const double d1 = s_h[0];
        if (d1 == 0.0) {
          s_st->reason = -5;
          s_st->done = 1;
        }
        sc[1] = sc[5] / d1;
      }
      bar_sync_named(FZ_BAR_ALL, nc);
      if (s_st->done) break;
      // ---- S = R - alpha V, T = M^-1 A S
      {
        const double alpha = sc[1];
        auto fs = [&](double r, double v) { return __dsub_rn(r, __dmul_rn(alpha, v)); };
        for (int i = (eb >> 1) + tid; i < (ee >> 1); i += nc) {
          const double2 r = __ldcg(reinterpret_cast<const double2 *>(vR) + i), v = __ldcg(reinterpret_cast<const double2 *>(vV) + i);
          reinterpret_cast<double2 *>(vS)[i] = make_double2(fs(r.x, v.x), fs(r.y, v.y));
        }
        if (tid == 0) {
          if (eb > e0) vS[e0] = fs(__ldcg(vR + e0), __ldcg(vV + e0));
          if (ee < e1 && ee >= eb) vS[e1 - 1] = fs(__ldcg(vR + e1 - 1), __ldcg(vV + e1 - 1));
        }
      }
      push(vS);
      FZ_STAMP(3);
      barrier_gpu();
      hseq++;
      FZ_STAMP(4);
      if (__ldcg(&a.bar[2])) break;
      sp_phase(1, vS, 1.0, nullptr, vT);
      FZ_STAMP(0);
      bar_sync_named(FZ_BAR_ALL, nc);
      dots(vT, vS, 0, 1, s_h, true);  // (T, S), (T, T)
      FZ_STAMP(1);
      reduce_numbers(2, false);
      FZ_STAMP(2);
      if (__ldcg(&a.bar[2])) break;
      if (tid == 0) {
        const double d1 = s_h[0], d2 = s_h[1];
        sc[6] = d2 == 0.0 ? 1.0 : 0.0;
        sc[4] = d2 == 0.0 ? 0.0 : d1 / d2;
      }
      bar_sync_named(FZ_BAR_ALL, nc);
      // ---- x += alpha P + omega S, R = S - omega T   (T . T == 0: x += alpha P and the solve ends, KSP_CONVERGED_ATOL)
      {
        const double alpha = sc[1], omega = sc[4];
        const bool t0 = sc[6] != 0.0;
        auto fx = [&](double xe, double pe, double se) {
          return t0 ? __dadd_rn(xe, __dmul_rn(alpha, pe)) : __dadd_rn(xe, __dadd_rn(__dmul_rn(alpha, pe), __dmul_rn(omega, se)));
        };
        auto fr = [&](double se, double te) { return __dsub_rn(se, __dmul_rn(omega, te)); };
        for (int i = (eb >> 1) + tid; i < (ee >> 1); i += nc) {
          const double2 pe = __ldcg(reinterpret_cast<const double2 *>(vP) + i), se = __ldcg(reinterpret_cast<const double2 *>(vS) + i),
                        xe = __ldcg(reinterpret_cast<const double2 *>(a.xp) + i);
          reinterpret_cast<double2 *>(a.xp)[i] = make_double2(fx(xe.x, pe.x, se.x), fx(xe.y, pe.y, se.y));
          if (!t0) {
            const double2 te = __ldcg(reinterpret_cast<const double2 *>(vT) + i);
            reinterpret_cast<double2 *>(vR)[i] = make_double2(fr(se.x, te.x), fr(se.y, te.y));
          }
        }
        if (tid == 0) {
          if (eb > e0) {
            a.xp[e0] = fx(__ldcg(a.xp + e0), __ldcg(vP + e0), __ldcg(vS + e0));
            if (!t0) vR[e0] = fr(__ldcg(vS + e0), __ldcg(vT + e0));
          }
          if (ee < e1 && ee >= eb) {
            const int q = e1 - 1;
            a.xp[q] = fx(__ldcg(a.xp + q), __ldcg(vP + q), __ldcg(vS + q));
            if (!t0) vR[q] = fr(__ldcg(vS + q), __ldcg(vT + q));
          }
        }
      }
      FZ_STAMP(3);
      bar_sync_named(FZ_BAR_ALL, nc);
      dots(vR, vRP, 0, 1, s_h, true);  // (R, RP), (R, R)
      FZ_STAMP(1);
      reduce_numbers(2, false);
      FZ_STAMP(2);
      if (__ldcg(&a.bar[2])) break;
      if (tid == 0) {
        s_st->its += 1;
        if (sc[6] != 0.0) {
          s_st->res = 0.0;
          s_st->reason = 3;
          s_st->done = 1;
        } else {
          const double dp = sqrt(s_h[1]);
          s_st->res = dp;
          sc[0] = sc[5];
          sc[2] = sc[4];
          int reason = 0;
          const double ttol = fmax(u.rtol * s_st->rnorm0, u.atol);
          if (dp != dp) reason = -9;
          else if (dp <= ttol) reason = (dp < u.atol) ? 3 : 2;
          else if (dp >= u.dtol * s_st->rnorm0) reason = -4;
          if (!reason && s_st->its >= u.maxit) reason = -3;
          if (reason) {
            s_st->reason = reason;
            s_st->done = 1;
          }
        }
      }
      bar_sync_named(FZ_BAR_ALL, nc);
      if (profiler) s_prof[6]++;
    }
  } else {
  double *w_old = a.wa, *w_new = a.wb;
  bool first = true, done = false;
  if (tid == 0) {
    s_st->res = 0.0; s_st->rnorm0 = 0.0; s_st->scal1 = 1.0;
    s_st->its = 0; s_st->it_inner = 0; s_st->reason = 0; s_st->done = 0;
  }
  FZ_STAMP(5);
  while (true) {
    // ================================ start of a restart cycle: w = M^-1 (b - A x), res = |w|
    sp_phase(first ? 0 : 2, a.xp, 1.0, nullptr, w_new);
    FZ_STAMP(0);
    bar_sync_named(FZ_BAR_ALL, nc);
    push(w_new);
    dots(w_new, w_new, 0, 1, s_h, false);
    FZ_STAMP(1);
    fz_reduce_ll(a, R, 1, ++rseqB, 1, s_h, s_red, true, multi, bseq + 1, WB_P2P_SLOT_FB, 1, cta, tid, nc);
    if (tid == 0) {  // k_gmres_begin
      const double res = sqrt(s_h[0]);
      s_st->res = res;
      s_st->it_inner = 0;
      if (first) {
        s_st->rnorm0 = res;
        s_st->its = 0;
        s_st->reason = 0;
        int reason = 0;
        const double ttol = fmax(u.rtol * res, u.atol);
        if (res != res) reason = -9;
        else if (res <= ttol) reason = (res < u.atol) ? 3 : 2;
        if (!reason && res == 0.0) reason = 3;
        if (reason) {
          s_st->reason = reason;
          s_st->done = 1;
        }
      }
      s_st->scal1 = res > 0.0 ? 1.0 / res : 1.0;
      s_rs[0] = res;
      s_pn[2] = res;  // residual norm at the start of this restart cycle
    }
    bar_sync_named(FZ_BAR_ALL, nc);
    bseq++;
    hseq++;
    FZ_STAMP(2);
    if (__ldcg(&a.bar[2])) break;
    done = s_st->done != 0;
    if (done) break;
    {
      double *t = w_old;
      w_old = w_new;
      w_new = t;
    }
    // ================================ Arnoldi iterations of the cycle
    int it = 0;
    while (true) {
      sp_phase(1, w_old, s_st->scal1, a.V + (size_t)it * a.ld, w_new);
      FZ_STAMP(0);
      bar_sync_named(FZ_BAR_ALL, nc);
      dots(w_new, a.V, a.ld, it + 1, s_h, pyth);
      FZ_STAMP(1);
      fz_reduce_ll(a, R, 0, ++rseqA, it + 1 + (pyth ? 1 : 0), s_h, s_red, false, multi, aseq + 1, WB_P2P_SLOT_FA, WB_P2P_MAXV, cta,
                   tid, nc);
      aseq++;
      FZ_STAMP(2);
      if (__ldcg(&a.bar[2])) break;
      if (tid <= it) s_cf[tid] = -s_h[tid];
      if (pyth && tid == 0) {
        // |w - sum h_j v_j|^2 = |w|^2 - sum h_j^2 for an orthonormal basis.  Every CTA of every rank evaluates this from
        // the same reduced numbers with the same instructions, so all take the same branch.  When the difference
        // cancels below FZ_PYTH_MIN of |w|^2, or the estimated loss of orthogonality of the basis would change it by
        // more than FZ_PYTH_ORTH, the norm is measured explicitly instead (second reduction, as VecNorm).
        const double ww = s_h[it + 1];
        double sh = 0.0;
        for (int j = 0; j <= it; j++) sh += s_h[j] * s_h[j];
        const double n2 = ww - sh;
        // the identity needs an orthonormal basis: classical Gram-Schmidt loses orthogonality like
        // eps * (residual at the start of the cycle / residual now)^2, which enters n2 as that times |w|^2
        const double red = s_pn[2] / fmax(s_st->res, 1e-300);
        s_pn[0] = n2;
        s_pn[1] = (n2 > FZ_PYTH_MIN * ww && 2.3e-16 * red * red * ww < FZ_PYTH_ORTH * n2) ? 1.0 : 0.0;
      }
      bar_sync_named(FZ_BAR_ALL, nc);
      const bool fast = pyth && s_pn[1] != 0.0;
      maxpy(w_new, a.V, a.ld, it + 1, fast ? nullptr : s_cf);
      push(w_new);
      FZ_STAMP(3);
      if (fast) {
        // no number to exchange: a barrier of this GPU's CTAs orders the rows of w before the next product reads them;
        // the neighbours' rows arrive through the flagged halo entries
        fz_reduce_ll(a, R, 1, ++rseqB, 1, s_cf, s_red, true, false, 0, 0, 1, cta, tid, nc);
      } else {
        fz_reduce_ll(a, R, 1, ++rseqB, 1, s_cf, s_red, true, multi, bseq + 1, WB_P2P_SLOT_FB, 1, cta, tid, nc);
        bseq++;
      }
      if (tid == 0) {
        fz_gmres_update(u, s_st, fast ? s_pn[0] : s_cf[0], s_h, s_H, s_cs, s_sn, s_rs);
        // end of the cycle (restart length reached or finished): back substitution y = H^-1 rs (k_gmres_solve_y)
        const int itn = s_st->it_inner;
        if (s_st->done || itn >= m) {
          for (int k = itn - 1; k >= 0; k--) {
            double sy = s_rs[k];
            for (int j = k + 1; j < itn; j++) sy -= s_H[(size_t)(m + 1) * j + k] * s_cf[j];
            s_cf[k] = sy / s_H[(size_t)(m + 1) * k + k];
          }
        }
      }
      bar_sync_named(FZ_BAR_ALL, nc);
      hseq++;
      FZ_STAMP(4);
      if (profiler) s_prof[6]++;
      if (__ldcg(&a.bar[2])) break;
      {
        double *t = w_old;
        w_old = w_new;
        w_new = t;
      }
      it = s_st->it_inner;
      done = s_st->done != 0;
      if (done || it >= m) break;
    }
    if (__ldcg(&a.bar[2])) break;
    // ================================ x += sum_j y_j V_j over the columns built in this cycle (y is in s_cf)
    if (it > 0) maxpy(a.xp, a.V, a.ld, it, nullptr);
    if (done) break;
    // another cycle follows: its residual needs everybody's x (the neighbours' on the boundary)
    push(a.xp);
    if (tid == 0) s_cf[0] = 0.0;
    fz_reduce_ll(a, R, 1, ++rseqB, 1, s_cf, s_red, true, multi, bseq + 1, WB_P2P_SLOT_FB, 1, cta, tid, nc);  // barrier only
    bseq++;
    hseq++;
    first = false;
    FZ_STAMP(5);
    if (__ldcg(&a.bar[2])) break;
  }
  }
  // ---- the solution in the caller's ordering; the solver state for the host; stop the producer
  bar_sync_named(FZ_BAR_ALL, nc);
  for (int r = R0 + tid; r < R1; r += nc) {
    const size_t ob = (size_t)a.perm[r] * BS;
#pragma unroll
    for (int i = 0; i < BS; i++) a.xout[ob + i] = a.xp[(size_t)r * BS + i];
  }
  if (tid == 0) *s_stop = 1;
  if (cta == 0 && tid == 0) {
    KspState *st = u.st;
    st->res = s_st->res;
    st->rnorm0 = s_st->rnorm0;
    st->its = s_st->its;
    st->reason = s_st->reason;
    st->it_inner = s_st->it_inner;
    *u.done = s_st->done;
    if (multi) {
      a.fseq[0] = hseq;
      a.fseq[1] = aseq;
      a.fseq[2] = bseq;
    }
  }
  FZ_STAMP(5);
  if (profiler)
    for (int k = 0; k < 16; k++) a.prof[(size_t)cta * 16 + k] += s_prof[k];
#undef FZ_STAMP
}

// ================================================================ host: the solve

bool wb_fused_usable(const wb_mat *A, const wb_pc *pc, const wb_ksp_opts *o) {
  if (!pc || !pc->fused || !pc->blocked || pc->type != WB_PC_BJACOBI_ILU0 || !fused_mode()) return false;
  if (o->type != WB_KSP_GMRES && o->type != WB_KSP_BCGS) return false;
  if (o->type == WB_KSP_BCGS) {
    static int on = -1;  // WB_FUSED_BCGS=0: BiCGStab stays with the launch-per-operation kernels
    if (on < 0) {
      const char *e = getenv("WB_FUSED_BCGS");
      on = e ? atoi(e) : 1;
    }
    if (!on) return false;
  }
  const int m = o->restart > 0 ? o->restart : 30;
  if (o->type == WB_KSP_GMRES && m + 1 > KRY_MAXV) return false;
  // 3x3 blocks leave the persistent kernel with 6 consumer warps per SM: with several sub-domains per CTA (the
  // bandwidth-bound regime) the launch-per-operation kernels stream faster (measured on config 4: 294 vs 274 us per
  // iteration); with one or two sub-domains per CTA (latency-bound) the persistent kernel wins (config 5: 145 vs 177)
  if (fused_mode() == 1 && pc->bs >= 3 && pc->fused->sd_cap > 2) return false;  // WB_FUSED=2 forces it
  const wb_ctx *c = A->ctx;
  if (c->nranks > 1) return c->p2p.on && A == &c->J && c->nranks <= WB_P2P_MAX_RANKS;
  return A->ncolb == A->nb;
}

// boundary rows of every CTA and where they go in the neighbours' ghost buffers
static int build_push_lists(wb_ctx *c, wb_pc *pc) {
  WbFusedPlan *f = pc->fused;
  if (f->push_built) return 0;
  const WbHalo &h = c->halo;
  const WbP2P &p = c->p2p;
  std::vector<int32_t> idx(std::max(h.nsend, 1));
  if (h.nsend > 0) WB_CUDA(wb_memcpy_sync(idx.data(), h.d_send_idx, sizeof(int32_t) * h.nsend, cudaMemcpyDeviceToHost));
  struct E { int row, rank, off; };
  std::vector<std::vector<E>> per(f->ncta);
  // CTA of a row in the sub-domain-major ordering
  std::vector<int> row_end(f->ncta);
  for (int ct = 0; ct < f->ncta; ct++) {
    const int4 C = f->h_cta[ct];
    const int4 last = pc->h_blk[C.x + C.y - 1];
    row_end[ct] = last.x + last.y;
  }
  for (int n = 0; n < h.nneigh; n++)
    for (int k = h.send_ptr[n]; k < h.send_ptr[n + 1]; k++) {
      const int row = f->h_invperm[idx[k]];
      const int ct = (int)(std::upper_bound(row_end.begin(), row_end.end(), row) - row_end.begin());
      WB_CHECK(ct < f->ncta, "fused GMRES: boundary row outside every CTA");
      per[ct].push_back({row, h.rank[n], p.peer_off[n] + (k - h.send_ptr[n])});
    }
  std::vector<int32_t> ptr(f->ncta + 1, 0), row, rank, off;
  for (int ct = 0; ct < f->ncta; ct++) {
    for (const E &e : per[ct]) {
      row.push_back(e.row);
      rank.push_back(e.rank);
      off.push_back(e.off);
    }
    ptr[ct + 1] = (int32_t)row.size();
  }
  WB_TRY(up(&f->d_push_ptr, ptr));
  WB_TRY(up(&f->d_push_row, row));
  WB_TRY(up(&f->d_push_rank, rank));
  WB_TRY(up(&f->d_push_off, off));
  f->push_built = true;
  return 0;
}

int wb_gmres_fused(wb_mat *A, wb_pc *pc, const wb_ksp_opts *o, const double *d_b, double *d_x, int *its, int *reason,
                   double *rnorm) {
  wb_ctx *c = A->ctx;
  WbFusedPlan *f = pc->fused;
  const size_t n = (size_t)A->nb * A->bs;
  const bool bcgs = o->type == WB_KSP_BCGS;
  const int m = bcgs ? 30 : (o->restart > 0 ? o->restart : 30);  // BiCGStab keeps its six vectors in the basis storage
  KspWork *wp;
  WB_TRY(wb_ensure_work(c, n, m, &wp));
  KspWork &w = *wp;
  double *hcol = w.small, *H = hcol + (m + 2), *cs = H + (size_t)(m + 1) * m, *sn = cs + (m + 1), *rs = sn + (m + 1),
         *yv = rs + (m + 2), *scal = yv + (m + 1);
  const bool multi = c->nranks > 1;
  if (multi) WB_TRY(build_push_lists(c, pc));
  WB_TRY(wb_fused_refresh(pc));  // the operator of this solve is the matrix as it is now
  WB_CUDA(cudaMemsetAsync(w.d_done, 0, sizeof(int), c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_st, 0, sizeof(KspState), c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_bar, 0, 64 * sizeof(int), c->stream));
  WB_CUDA(cudaMemsetAsync(w.d_ll, 0, WB_LL_BYTES, c->stream));
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.cta = f->d_cta; a.sd = f->d_sd; a.blk = pc->d_blk; a.slice = f->d_slice; a.recA = f->d_recA; a.recB = f->d_recB;
  a.rec_ptr = f->d_rec_ptr; a.perm = pc->d_blk_rows; a.sell = f->d_sell; a.stream = pc->d_stream;
  a.ng = f->ng; a.gt = f->gt; a.nc = f->nc; a.lt = f->lt; a.off_gm = f->off_gm; a.off_prof = f->off_prof; a.sd_cap = f->sd_cap; a.slice_cap = f->slice_cap; a.rec_cap = f->rec_cap;
  a.zs_words = f->zs_words; a.ring_bytes = f->ring_bytes;
  a.off_red = f->off_red; a.off_misc = f->off_misc; a.off_sd = f->off_sd; a.off_slice = f->off_slice;
  a.off_rec = f->off_rec; a.off_zs = f->off_zs; a.off_ring = f->off_ring;
  a.nb = A->nb; a.ncta = f->ncta; a.ld = w.ld;
  a.b = d_b; a.xout = d_x; a.V = w.V; a.wa = w.tmp; a.wb = w.tmp + w.ld; a.xp = w.tmp + 2 * w.ld; a.llpart = w.d_ll;
  a.bar = w.d_bar;
  a.upd = {hcol, H, cs, sn, rs, scal, w.d_st, w.d_done, o->rtol, o->atol, o->dtol, m, o->maxit};
  a.yv = yv;
  a.prof = w.d_prof;
  if (g_fused_norm < 0) {
    const char *e = getenv("WB_FUSED_NORM");
    g_fused_norm = e ? atoi(e) : 1;
  }
  a.pyth = (g_fused_norm == 2 || (g_fused_norm == 1 && multi)) ? 1 : 0;
  {
    static int jitter = -1;
    if (jitter < 0) {
      const char *e = getenv("WB_FUSED_JITTER");
      jitter = e ? atoi(e) : 0;
    }
    a.jitter = jitter;
  }
  if (multi) {
    a.P = c->p2p.dev;
    a.nneigh = c->halo.nneigh;
    a.nb_rank = c->p2p.d_nb_rank;
    for (int r = 0; r < WB_P2P_MAX_RANKS; r++) {
      a.ll_off[r] = r < c->nranks ? c->p2p.peer_ll_off[r] : 0;
      a.ll_stride[r] = r < c->nranks ? c->p2p.peer_ll_stride[r] : 0;
    }
    a.push_ptr = f->d_push_ptr; a.push_row = f->d_push_row; a.push_rank = f->d_push_rank; a.push_off = f->d_push_off;
    a.fseq = c->p2p.d_fseq;
  }
  const int threads = f->nc + 32;  // consumers + the producer warp
  void *args[] = {&a};
  cudaError_t e;
#define FZ_LAUNCH(BS_, BC_)                                                                                          \
  do {                                                                                                             \
    WB_CUDA(cudaFuncSetAttribute(k_gmres_fused<BS_, BC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem)); \
    e = cudaLaunchCooperativeKernel((void *)k_gmres_fused<BS_, BC_>, dim3(f->ncta), dim3(threads), args, f->smem,  \
                                    c->stream);                                                                    \
  } while (0)
  switch (A->bs * 2 + (bcgs ? 1 : 0)) {
    case 2: FZ_LAUNCH(1, 0); break;
    case 3: FZ_LAUNCH(1, 1); break;
    case 4: FZ_LAUNCH(2, 0); break;
    case 5: FZ_LAUNCH(2, 1); break;
    case 6: FZ_LAUNCH(3, 0); break;
    default: FZ_LAUNCH(3, 1); break;
  }
#undef FZ_LAUNCH
  WB_CHECK(e == cudaSuccess, "fused GMRES: cooperative launch failed: %s", cudaGetErrorString(e));
  WB_LAUNCH(c);
  WB_TRY(wb_fetch_state(w));
  int h_bar[16] = {0};
  WB_CUDA(wb_memcpy_sync(h_bar, w.d_bar, sizeof(h_bar), cudaMemcpyDeviceToHost));
  WB_CHECK(!h_bar[2],
           "fused GMRES: a wait timed out (a CTA or a peer GPU is missing) [%s, CTA %d, thread %d, record %d; %d groups, "
           "%d records per group]",
           h_bar[8] == 1 ? "waiting for a record's bulk copy" : "in a reduction", h_bar[9], h_bar[10], h_bar[11], f->ng, f->rec_cap);
  *its = w.h_st->its;
  *reason = w.h_st->reason;
  *rnorm = w.h_st->res;
  return 0;
}

// per-phase device time of CTA 0 accumulated over the fused solves of this context since the last reset
// (nanoseconds): SpMV + PC, dots, dots barrier, multi-AXPY + norm, norm barrier, other; iterations
extern "C" int wb_ksp_fused_profile(wb_ctx *c, double *ns7, int reset) {
  KspWork *wp = wb_find_work(c);
  for (int k = 0; k < 7; k++) ns7[k] = 0.0;
  if (!wp) return 0;
  WB_CUDA(cudaSetDevice(c->device));
  unsigned long long h[16];
  WB_CUDA(wb_memcpy_sync(h, wp->d_prof, sizeof(h), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 7; k++) ns7[k] = (double)h[k];
  if (reset) WB_CUDA(wb_memset_sync(wp->d_prof, 0, WB_PROF_WORDS * sizeof(unsigned long long)));
  return 0;
}

// tuning aid (not part of the public header): the counters of every CTA, out[cta * 16 + k] (k < 8 as above; 9..11:
// SM cycles of thread 0 waiting for the first level record, in its own part of the levels, at the level barriers);
// returns the CTA count
extern "C" int wb_debug_fused_profile_all(wb_ctx *c, double *out, int max_ctas) {
  KspWork *wp = wb_find_work(c);
  if (!wp) return 0;
  WB_CUDA(cudaSetDevice(c->device));
  std::vector<unsigned long long> h(WB_PROF_WORDS);
  WB_CUDA(wb_memcpy_sync(h.data(), wp->d_prof, WB_PROF_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  const int n = std::min(max_ctas, WB_PROF_WORDS / 16);
  for (int k = 0; k < n * 16; k++) out[k] = (double)h[k];
  return n;
}
