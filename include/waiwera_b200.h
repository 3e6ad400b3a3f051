/*
 * waiwera_b200.h -- C ABI of the B200-native Newton-step engine for Waiwera.
 *
 * One opaque context per GPU (one process per GPU).  Every entry point takes
 * plain pointers and sizes; array arguments may be HOST or DEVICE pointers
 * (detected with cudaPointerGetAttributes; host arrays are staged through
 * pinned buffers).  Array record layouts are the reference's own, so buffers
 * obtained with VecGetArrayF90 on the Fortran side pass straight through:
 *   fluid record   7+nc-1 + nph*(8+nc-1) doubles   (src/fluid.F90:232-267)
 *   rock record    8 doubles                        (src/rock.F90:97-112)
 *   cell geometry  4 doubles                        (src/cell.F90:54-57)
 *   face geometry  12 doubles                       (src/face.F90:67-76)
 *   primaries      np doubles per cell, SCALED      (src/eos.F90:186-210)
 *   Jacobian       BAIJ: rowptr/colidx + bs*bs column-major blocks
 *                                                   (src/ode.F90:266-287)
 *
 * Set-up arrays that are index lists or per-source / per-boundary parameters (wb_set_boundaries, wb_set_sources,
 * wb_set_source_components, wb_set_source_controls, wb_set_halo, wb_set_pc_blocks) are read on the HOST: device
 * pointers are rejected there.
 *
 * Return value of every function: 0 ok; >0 recoverable physics / domain
 * error (the reference's `err` argument, already reduced over all GPUs of the
 * communicator -- src/mpi_utils.F90:46); <0 fatal (CUDA/NCCL/usage), message
 * via wb_last_error().  Calls are stream-ordered on the context's stream and
 * synchronous at return.  Not re-entrant per context.
 *
 * Each entry point cites the reference interface it stands in for.
 */
#ifndef WAIWERA_B200_H
#define WAIWERA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB_VERSION 100

/* thermodynamic formulation (src/thermodynamics_setup.F90:34-35) */
#define WB_THERMO_IAPWS 0
#define WB_THERMO_IFC67 1

/* equation of state (src/eos_setup.F90) */
#define WB_EOS_WE 0 /* water + energy, 2 primaries        (src/eos_we.F90)  */
#define WB_EOS_W 1  /* isothermal water, 1 primary        (src/eos_w.F90)   */
#define WB_EOS_WCE 2 /* water + CO2 + energy, 3 primaries (src/eos_wge.F90, src/eos_wce.F90,
                        src/ncg_co2_thermodynamics.F90) */

/* relative permeability curves (src/relative_permeability.F90:197-558) */
#define WB_RP_FULLY_MOBILE 0
#define WB_RP_LINEAR 1
#define WB_RP_PICKENS 2
#define WB_RP_COREY 3
#define WB_RP_GRANT 4
#define WB_RP_VAN_GENUCHTEN 5
#define WB_RP_TABLE 6
/* capillary pressure curves (src/capillary_pressure.F90:159-358) */
#define WB_CP_ZERO 0
#define WB_CP_LINEAR 1
#define WB_CP_VAN_GENUCHTEN 2
#define WB_CP_TABLE 3

#define WB_EOS_WAE 3 /* water + air + energy: eos_wge with the air NCG (src/eos_wae.F90,
                        src/ncg_air_thermodynamics.F90); 3 primaries, same kernels as WB_EOS_WCE */

#define WB_MAX_TABLE 16
#define WB_MAX_NP 3

typedef struct {
  int type;
  /* linear: liquid limits p[0..1], vapour limits p[2..3]; pickens: p[0]=power;
     corey/grant: p[0]=slr, p[1]=ssr;
     van Genuchten: p[0]=lambda, p[1]=slr, p[2]=sls, p[3]=sum_unity(0/1), p[4]=ssr;
     table: nl / nv points (saturation, value) sorted by saturation */
  double p[8];
  int nl, nv;
  double lx[WB_MAX_TABLE], ly[WB_MAX_TABLE], vx[WB_MAX_TABLE], vy[WB_MAX_TABLE];
} wb_relperm;

typedef struct {
  int type;
  /* linear: p[0..1] saturation limits, p[2] pressure;
     van Genuchten: p[0]=P0, p[1]=lambda, p[2]=slr, p[3]=sls, p[4]=Pmax, p[5]=apply_Pmax(0/1);
     table: n points sorted by saturation */
  double p[8];
  int n;
  double x[WB_MAX_TABLE], y[WB_MAX_TABLE];
} wb_cappress;

/* what flow_simulation_init reads from the JSON input for this path
   (src/flow_simulation.F90:882-1045, src/eos_we.F90:75-109) */
typedef struct {
  int eos;         /* WB_EOS_* */
  int thermo;      /* WB_THERMO_* */
  int extrapolate; /* thermodynamics.extrapolate */
  double pressure_scale, temperature_scale; /* eos.primary.scale.* (<=0: defaults 1e6, 1e2) */
  double partial_pressure_scale;            /* eos.primary.scale.partial_pressure for eos_wce; <= 0: adaptive
                                               scaling by the cell's total pressure, the reference default
                                               (src/eos_wge.F90:96-100, 639-674) */
  double eos_w_temperature;                 /* eos.temperature for eos_w */
  wb_relperm relperm;
  wb_cappress cappress;
  double gravity[3];
} wb_params;

typedef struct wb_ctx wb_ctx;

const char *wb_last_error(void);
int wb_version(void);

/* ---- life cycle ------------------------------------------------------- */
/* flow_simulation_init (src/flow_simulation.F90:882) / destroy (:1049) */
int wb_create(const wb_params *prm, int device, wb_ctx **out);
int wb_destroy(wb_ctx *ctx);
int wb_num_primary(const wb_ctx *ctx);
int wb_fluid_dof(const wb_ctx *ctx);

/*
 * Mesh arrays the path consumes.  Local cell numbering:
 * owned cells [0,nowned), partition ghost cells [nowned,ninterior), boundary
 * (Dirichlet) ghost cells [ninterior,ncell).  face_cells: 2 local cell indices
 * per flux face in mesh%flux_face order (src/mesh.F90:769-802), normal from
 * cell 1 to cell 2.  Also builds the BAIJ pattern of the Jacobian
 * (src/dm_utils.F90:1041-1051: row i = {i} U face neighbours with dofs; columns
 * in local numbering, ghost columns >= nowned).
 */
int wb_set_mesh(wb_ctx *ctx, int ncell, int ninterior, int nowned, int nface, const int32_t *face_cells,
                const double *face_geom, const double *cell_geom, const double *rock);

/* Jacobian pattern built by wb_set_mesh (DMCreateMatrix on interior_dm, src/ode.F90:275-283).
   Pointers are DEVICE pointers owned by the context (vals: nnzb*bs*bs). */
int wb_jacobian_pattern(wb_ctx *ctx, int *nb, int *bs, int *nnzb, const int32_t **rowptr, const int32_t **colidx,
                        double **vals);
/* copy pattern / values to host arrays (any may be NULL) */
int wb_jacobian_get(wb_ctx *ctx, int32_t *rowptr, int32_t *colidx, double *vals);
/* The cell -> face gather lists the residual / Jacobian kernels walk (the order in which the reference's face loop
   scatters into a cell, src/flow_simulation.F90:1410-1458): for owned cell i, entries cf_ptr[i] .. cf_ptr[i+1]-1 in
   ascending face order; cf_face = 2*face + side (0: the cell is cell 1 of the face), cf_other = the cell on the other
   side.  *ncf receives the number of entries; the arrays (host, any may be NULL) receive nowned+1 / ncf / ncf values. */
int wb_cell_faces_get(wb_ctx *ctx, int *ncf, int32_t *cf_ptr, int32_t *cf_face, int32_t *cf_other);

/* ---- multi-GPU -------------------------------------------------------- */
/* NCCL communicator over the ranks of the partition (replaces PetscSF /
   VecScatter, src/dm_utils.F90:480-498).  id: 128-byte ncclUniqueId made by
   wb_comm_unique_id on rank 0 and broadcast by the host (torch.distributed / MPI). */
int wb_comm_unique_id(void *id128);
int wb_comm_init(wb_ctx *ctx, int rank, int nranks, const void *id128);
/* halo plan: for each neighbour rank, the owned cells to send and the ghost cells
   (local indices in [nowned,ninterior)) to receive, CSR style. */
int wb_set_halo(wb_ctx *ctx, int nneigh, const int32_t *neigh_rank, const int32_t *send_ptr,
                const int32_t *send_idx, const int32_t *recv_ptr, const int32_t *recv_idx);
/* NVLink / NVSwitch peer-to-peer exchange for the three per-iteration exchanges of GMRES (ghost entries of x inside
   the SpMV, Gram-Schmidt coefficients, norm): peers write into CUDA-IPC-mapped buffers and publish sequence
   numbers, so none of them is a separate collective launch (replaces VecScatter + MPI_Allreduce inside
   KSPSolve_GMRES for the context's Jacobian).  Call after wb_comm_init + wb_set_halo: every rank exports a blob of
   wb_comm_p2p_blob_size() bytes, the host all-gathers them in rank order (torch.distributed / MPI), every rank
   opens the gathered array.  NCCL stays in use for everything else; on failure the context keeps using NCCL. */
int wb_comm_p2p_blob_size(void);
int wb_comm_p2p_export(wb_ctx *ctx, void *blob);
int wb_comm_p2p_open(wb_ctx *ctx, const void *blobs);
int wb_comm_p2p_enabled(const wb_ctx *ctx);
int wb_comm_p2p_disable(wb_ctx *ctx);
/* global offset of this rank's first owned cell (VecGetOwnershipRange / bs) */
int wb_set_global_offset(wb_ctx *ctx, int64_t first_cell, int64_t ncell_global);

/* ---- state ------------------------------------------------------------ */
/* fluid_init (src/flow_simulation.F90:2171-2287): regions + fluid properties from
   scaled primaries y[nowned*np] and region[nowned]. */
int wb_fluid_init(wb_ctx *ctx, const double *y, const int32_t *region);
/* Dirichlet boundary ghost cell (src/mesh.F90:1185-1202): rock copied from the interior
   cell, fluid record from UNSCALED primary and region. */
int wb_set_boundary(wb_ctx *ctx, int ghost_cell, int interior_cell, const double *primary, int region);
/* the same for n boundary cells at once: primary[n*np], region[n] */
int wb_set_boundaries(wb_ctx *ctx, int n, const int32_t *ghost_cells, const int32_t *interior_cells,
                      const double *primary, const int32_t *region);
/* Rock records of the interior cells replaced between time steps: rock[8*ninterior] (host array), the layout of
   wb_set_mesh.  Replaces flow_simulation_update_rock_properties (src/flow_simulation.F90:2051-2089), which
   pre_try_timestep calls (:2040-2047, src/timestepper.F90:2333) to apply the permeability / porosity tables of
   src/rock_control.F90:49-116 to the rock vector.  Boundary ghost cells keep the records copied at set-up
   (src/mesh.F90:1189-1193; the reference's controls list interior cells only, src/rock_setup.F90:404-412).  The
   per-face permeabilities are recomputed; fluid state, stored balances and the Jacobian pattern are untouched. */
int wb_set_rock(wb_ctx *ctx, const double *rock);
/* Fixed-rate sources / sinks (source%update_flow src/source.F90:375-480 and
   source_network%assemble_cell_inflows src/source_network.F90:296-355, called from cell_inflows
   src/flow_simulation.F90:1468-1473): rhs_i += flow / V_i.  cell: local owned cell; component: 1-based
   mass component, np = heat, 0 = all mass components (production only); rate > 0 injects `rate` of that
   component with specific enthalpy `enthalpy`, rate < 0 produces by mobility-weighted phase flow
   fractions (src/fluid.F90:374-456).  n = 0 removes all sources. */
int wb_set_sources(wb_ctx *ctx, int n, const int32_t *cell, const int32_t *component, const double *rate,
                   const double *enthalpy);
/* Injection and production component of every source of the last wb_set_sources (n = their number), as
   get_components reads them from the input (src/source_setup.F90:2052-2083): the reference picks one or the other from
   the sign of the CURRENT rate at every update (src/source.F90:372-380, 469-476), which matters for sources whose rate
   changes sign (rate tables, deliverability with direction "both").  Without this call the `component` of
   wb_set_sources applies to both signs.  injection_component: 1..np; production_component: 0..np (0 = all mass
   components by flow fraction). */
int wb_set_source_components(wb_ctx *ctx, int n, const int32_t *injection_component, const int32_t *production_component);
/* Source controls for n of the sources of the last wb_set_sources (source[k]: index into those arrays), re-evaluated
   from the fluid state of the source's cell at EVERY function evaluation -- the reference's source_network%update
   inside cell_inflows (src/flow_simulation.F90:1468-1473, src/source_network.F90:90-292) -- so that the
   finite-difference Jacobian carries their pressure / mobility dependence:
     deliverability  rate = -productivity * sum_p mobility_p * (P - reference_pressure) over the phases present
                     (deliverability_source_control_flow_rate, src/source_control.F90:359-403; productivity <= 0:
                     the fixed rate of wb_set_sources is kept)
     direction       0 both, 1 production only, 2 injection only (direction_source_control_iterator, :596-620)
     limit           "total" flow limiter: |rate| scaled down to limit (src/source_network_node.F90:245-315; <= 0: none)
   direction and limit may be NULL.  n = 0 removes all controls; wb_set_sources also removes them. */
int wb_set_source_controls(wb_ctx *ctx, int n, const int32_t *source, const double *productivity,
                           const double *reference_pressure, const int32_t *direction, const double *limit);
/* Recharge / injectivity controls (source input "recharge" / "injectivity": {"coefficient": c, "pressure": p};
   recharge_source_control_iterator, src/source_control.F90:554-577): rate = -c (P - p) with P the pressure of the
   source's cell, then the direction control and the limiters as for any source.  Edits the entries of these sources:
   call after wb_set_source_controls. */
int wb_set_source_recharge(wb_ctx *ctx, int n, const int32_t *source, const double *coefficient,
                           const double *reference_pressure);
/* rate of every source (order of wb_set_sources) for the state of the last unperturbed evaluation: the
   "rate" source output field */
int wb_get_source_rates(wb_ctx *ctx, double *rate);
/* Separators and limiters on the separated flows (src/separator.F90; source input "separator": {"pressure": p | [p1, p2]},
   "limiter": {"type": "water" | "steam", "limit": ..., "separator_pressure": ...} or {"total": ..., "water": ...,
   "steam": ...}, src/source_setup.F90:2255-2330, 3117-3276).  A producing source with a separator flashes its flow
   (mass rate and the flowing enthalpy of its cell's fluid, src/source_network.F90:197-216) through nstage[k] <= 2 stages
   at pressure[2 k], pressure[2 k + 1] (separator_separate, :212-260); limit_water / limit_steam (<= 0: none) limit the
   separated rates exactly as the total limiter limits the whole flow: one scale, the smallest over the limited flow types
   (source_network_node_limit_rate, src/source_network_node.F90:245-315).  Evaluated at every function evaluation like the
   other controls.  Call after wb_set_sources / wb_set_source_controls; n = 0 removes the separators. */
int wb_set_source_separators(wb_ctx *ctx, int n, const int32_t *source, const int32_t *nstage, const double *pressure,
                             const double *limit_water, const double *limit_steam);
/* Reference pressure of sources on deliverability as a table (source input "deliverability": {"pressure": {"enthalpy":
   [[h, P], ...]}} or {"pressure": [[P, Pref], ...]}; SRC_PRESSURE_TABLE_COORD_ENTHALPY / _PRESSURE in
   deliverability_source_control_flow_rate, src/source_control.F90:359-403): at every function evaluation the table of
   source[k] -- npts[k] <= WB_PRESSURE_TABLE_MAX points (x, y) at table[2 * WB_PRESSURE_TABLE_MAX * k ...], x increasing --
   is looked up at the flowing enthalpy of the source's cell (coordinate[k] = 0: phase enthalpies weighted by the flow
   fractions) or at its pressure (1), linearly or stepwise (step[k] != 0; null: linear), constant beyond the ends, and
   replaces the reference pressure of wb_set_source_controls.  Call after wb_set_sources; the tables stay until the source
   list changes; n = 0 removes them. */
#define WB_PRESSURE_TABLE_MAX 8
int wb_set_source_pressure_table(wb_ctx *ctx, int n, const int32_t *source, const int32_t *coordinate, const int32_t *step,
                                 const int32_t *npts, const double *table);
/* separator_stage_init (src/separator.F90:108-136): reference water and steam enthalpies of a stage at `pressure` */
int wb_separator_stage(wb_ctx *ctx, double pressure, double *ref_water_enthalpy, double *ref_steam_enthalpy);
/* the separated-flow source output fields [nsources][5]: water_rate, water_enthalpy, steam_rate, steam_enthalpy,
   steam_fraction (src/source_network_node.F90:95-112), state of the last unperturbed evaluation */
int wb_get_source_separated(wb_ctx *ctx, double *out5);
/* current fluid records of all local cells, reference AoS layout [ncell*fluid_dof] */
int wb_get_fluid(wb_ctx *ctx, double *fluid);
int wb_get_regions(wb_ctx *ctx, int32_t *region);
/* ode hooks pre_iteration / pre_timestep / pre_retry_timestep
   (src/flow_simulation.F90:2108, :2022, :2093) */
int wb_pre_iteration(wb_ctx *ctx);
int wb_pre_timestep(wb_ctx *ctx);
int wb_pre_retry_timestep(wb_ctx *ctx);

/* time-stepping method (context%residual, src/timestepper.F90:2223-2245): selects the residual that
   wb_residual_be, wb_jacobian_be(_colored) and wb_newton_solve_be evaluate from then on:
     WB_METHOD_BEULER   backwards_Euler_residual (:345-374)  r = L - L_last - dt R            (default)
     WB_METHOD_BDF2     BDF2_residual (:378-427)  r = (1+2q) L - (q+1)^2 L_last + q^2 L_last2 - dt (q+1) R,
                        q = dt / dt_last; lhs_last2 = L two steps back (host or device, nowned*np)
     WB_METHOD_DIRECTSS direct_ss_residual (:431-452)  r = R  (dt and lhs_last arguments are ignored) */
#define WB_METHOD_BEULER 0
#define WB_METHOD_BDF2 1
#define WB_METHOD_DIRECTSS 2
int wb_set_method(wb_ctx *ctx, int method, double dt_last, const double *lhs_last2);

/* ---- function evaluation (ode_type lhs / rhs, SNES_residual) ----------- */
/* pre_eval (src/flow_simulation.F90:2126): fluid_properties for y.  perturbed /
   nperturbed are the block columns MatFDColoring perturbed (src/dm_utils.F90:1544);
   nperturbed == 0 is the "unperturbed" evaluation that also updates the stored fluid. */
int wb_pre_eval(wb_ctx *ctx, const double *y, const int32_t *perturbed, int nperturbed);
/* cell_balances (src/flow_simulation.F90:1242) and cell_inflows (:1334) for the state of the
   last wb_pre_eval; out arrays nowned*np */
int wb_cell_balances(wb_ctx *ctx, double *lhs);
int wb_cell_inflows(wb_ctx *ctx, double *rhs);
/* SNES_residual + backwards_Euler_residual (src/timestepper.F90:587, :345):
   r = L(y) - lhs_last - dt*R(y); includes pre_eval.  lhs / rhs may be NULL. */
int wb_residual_be(wb_ctx *ctx, const double *y, const double *lhs_last, double dt, const int32_t *perturbed,
                   int nperturbed, double *lhs, double *rhs, double *r);
/* vec_max_pointwise_abs_scale (src/dm_utils.F90:644-685): max_i |v_i| / max(|scale_i|, tol)
   and its first (global) index, reduced over all ranks */
int wb_max_scaled(wb_ctx *ctx, const double *v, const double *scale, double tol, double *maxval,
                  int64_t *maxloc);

/* ---- Jacobian (SNESComputeJacobianDefaultColor, src/timestepper.F90:1584-1611) -- */
/* Local finite differences with the MATMFFD_DS step rule (err, umin): fills the
   context's BAIJ matrix with J(:,j) = (F(y + h_j e_j) - F(y)) / h_j for the BE residual.
   vals_out (host or device, may be NULL) receives nnzb*bs*bs values. */
int wb_jacobian_be(wb_ctx *ctx, const double *y, const double *lhs_last, double dt, double fd_err,
                   double fd_umin, double *vals_out);
/* the same through the reference's colouring loop (one masked residual evaluation per colour
   and per variable); slower, kept for parity checks against the local assembly */
int wb_jacobian_be_colored(wb_ctx *ctx, const double *y, const double *lhs_last, double dt, double fd_err,
                           double fd_umin, double *vals_out, int *ncolors);

/* ---- transitions (post_linesearch, src/flow_simulation.F90:2419-2576) --- */
int wb_fluid_transitions(wb_ctx *ctx, const double *y_old, double *search, double *y, int *changed_search,
                         int *changed_y);

/* ---- Mat / PC / KSP (PETSc plug-in seam, src/timestepper.F90:1645-1836) -- */
typedef struct wb_mat wb_mat;
typedef struct wb_pc wb_pc;

/* MatCreateBAIJ + MatSetValuesBlocked: square block matrix, nb block rows, ncolb block
   columns (>= nb; columns >= nb are ghost columns filled by the halo), column-major blocks. */
int wb_mat_create(wb_ctx *ctx, int nb, int ncolb, int bs, int nnzb, const int32_t *rowptr,
                  const int32_t *colidx, const double *vals, wb_mat **out);
int wb_mat_set_values(wb_mat *A, const double *vals);
/* MatSeqBAIJGetArray: copies the nnzb*bs*bs block values to vals (host or device) */
int wb_mat_get_values(wb_mat *A, double *vals);
int wb_mat_destroy(wb_mat *A);
/* the context's own Jacobian as a wb_mat (borrowed; do not destroy) */
int wb_jacobian_mat(wb_ctx *ctx, wb_mat **out);
/* MatMult (MatMult_SeqBAIJ_2/3, MatMult_MPIBAIJ incl. ghost scatter): y = A x; x, y: nb*bs */
int wb_mat_mult(wb_mat *A, const double *x, double *y);

#define WB_PC_NONE 0
#define WB_PC_PBJACOBI 1 /* point-block Jacobi */
#define WB_PC_BJACOBI_ILU0 2 /* block Jacobi, ILU(0) natural ordering on each block */
#define WB_PC_ASM_ILU0 3     /* PCASM (restricted, overlap 1) on the same sub-domains, ILU(0) on each extended sub-domain;
                                the overlap stays inside the rank (rows of other ranks are not fetched) */
/* PCSetUp.  nblocks: number of block-Jacobi sub-domains on this GPU (PETSc
   -pc_bjacobi_local_blocks; 1 = one ILU(0) over all owned rows, the PETSc default).
   block_of_row[nb] (host, may be NULL => contiguous equal split) assigns rows to blocks. */
int wb_pc_setup(wb_mat *A, int type, int nblocks, const int32_t *block_of_row, wb_pc **out);
/* PCSetUp after the matrix values changed (same pattern): numeric part only */
int wb_pc_refactor(wb_pc *pc);
int wb_pc_apply(wb_pc *pc, const double *r, double *z);
int wb_pc_destroy(wb_pc *pc);

#define WB_KSP_GMRES 0
#define WB_KSP_BCGS 1
typedef struct {
  int type, restart, maxit;
  double rtol, atol, dtol;
} wb_ksp_opts;
/* KSPSolve with zero initial guess, left preconditioning; reason follows KSPConvergedReason */
int wb_ksp_solve(wb_mat *A, wb_pc *pc, const wb_ksp_opts *opts, const double *b, double *x, int *its,
                 int *reason, double *rnorm);

/* Krylov iterations enqueued between host-side convergence checks (the kernels skip their
   work once the device-side flag says converged, so results do not depend on it); default 4 */
int wb_ksp_set_check_every(int k);
/* GMRES with block-Jacobi / ILU(0) sub-domains normally runs as one persistent kernel for the whole solve
   (sub-domain-resident: SpMV, PC apply and Gram-Schmidt of a sub-domain stay on one SM; restart <= 31); 0 selects the
   launch-per-operation solver instead, 2 uses the persistent kernel wherever it can run (1, the default, leaves 3x3
   block systems with many sub-domains per SM to the launch-per-operation kernels); also: environment WB_FUSED.
   Takes effect at the next PC set-up. */
int wb_ksp_set_fused(int on);
/* Norm of the new Krylov vector inside the persistent kernel.  PETSc's GMRES (classical Gram-Schmidt, no refinement)
   takes VecNorm of the orthogonalised vector: a second all-reduce per iteration.  Mode 0 does exactly that.  Mode 2
   takes the norm from the first reduction instead, |w|^2 - sum_j h_j^2 with |w|^2 accumulated on the dot-product pass
   (the "Pythagorean" form used by low-synchronisation GMRES variants), and falls back to the explicit norm whenever
   the difference is below 1e-4 |w|^2 or the estimated loss of orthogonality of the basis (eps times the square of the
   residual reduction inside the restart cycle) would change it by more than 1e-6; what is left of the second reduction is a barrier inside each GPU.  The value
   differs from VecNorm's by rounding times |w|^2 / |w'|^2 and by the loss of orthogonality of the basis, so iteration
   counts can move by a few in thousands; converged solutions agree to the solver tolerance.  Mode 1 (default) uses
   mode 2 only when the solve spans several GPUs, where it removes an NVLink all-gather from every iteration.
   Environment: WB_FUSED_NORM. */
int wb_ksp_set_fused_norm(int mode);

/* ---- Newton (SNESSolve as configured by timestepper.F90:1552-1641) ------ */
typedef struct {
  int max_iterations, min_iterations;
  double rel_tol, abs_tol, update_rel_tol, update_abs_tol;
  double fd_err, fd_umin;
  int pc_type, pc_nblocks;
  wb_ksp_opts ksp;
} wb_newton_opts;
typedef struct {
  int reason, iterations, linear_iterations;
  double max_residual[32];
  int lin_its[32];
  int lin_reason[32];    /* KSPConvergedReason of the linear solve of every Newton iteration */
  double lin_rnorm[32];  /* its final (preconditioned) residual norm */
} wb_newton_result;
/* sub-domain of every owned row for the block-Jacobi preconditioner the Newton solve sets up
   (arbitrary index sets as with PCASMSetLocalSubdomains at overlap 0; NULL restores the default
   contiguous split into opts->pc_nblocks ranges).  block_of_row: host array, nowned entries. */
int wb_set_pc_blocks(wb_ctx *ctx, const int32_t *block_of_row);
/* one backward-Euler step solve: y in/out (scaled primaries of owned cells) */
int wb_newton_solve_be(wb_ctx *ctx, const wb_newton_opts *opts, double dt, const double *lhs_last, double *y,
                       wb_newton_result *res);

/* ---- passive tracers: the auxiliary linear problem ------------------
   After a converged Newton solve the reference assembles and solves one LINEAR system for the tracer
   mass fractions (src/timestepper.F90:2347-2353): method%setup_linear (:458-581) builds
   A = Al - dt Ar, b = Al_last x_last + dt br from aux_lhs (src/flow_simulation.F90:1489-1556) and aux_rhs
   (:1560-1833: upstream advection with the stored phase fluxes, diffusion, production, injection, Arrhenius
   decay), aux_pre_solve (:1837-1959) pins absent phases and Dirichlet cells, KSPSolve solves it.  Here A_aux
   lives on the Jacobian's block pattern with bs = nt (tracers do not couple: diagonal blocks); rows of
   Dirichlet ghost cells are eliminated into b.  All calls use the fluid state of the last UNPERTURBED
   evaluation (wb_pre_eval / wb_residual_be with nperturbed = 0, or the end of wb_newton_solve_be), as the
   reference does.  The time-stepping method is the one set with wb_set_method. */
#define WB_MAX_TRACERS 3
/* setup_tracers (src/tracer.F90:64-150): phase[nt] 1-based phase index, diffusion[nt] (m2/s),
   decay[nt] constant (1/s), activation[nt] energy (J/mol); any of the last three may be NULL (zeros).
   nt = 0 removes the tracers. */
int wb_set_tracers(wb_ctx *ctx, int nt, const int32_t *phase, const double *diffusion, const double *decay,
                   const double *activation);
/* source%tracer_injection_rate (src/source.F90): rate[nsources*nt] (kg/s) in the order of the last
   wb_set_sources; NULL removes them.  Only injecting sources (rate > 0) use it. */
int wb_set_tracer_injection(wb_ctx *ctx, const double *rate);
/* aux_lhs (src/flow_simulation.F90:1489-1556): al[nowned*nt] = porosity * saturation * density of the tracer's phase */
int wb_tracer_cell_balances(wb_ctx *ctx, double *al);
/* setup_linear + aux_pre_solve.  al_last / x_last: balance coefficients and mass fractions at the last step
   (nowned*nt); al_last2 / x_last2: two steps back (BDF2 only, else NULL); x_boundary[(ncell-ninterior)*nt]:
   mass fractions of the Dirichlet ghost cells (NULL if there are none).  Outputs: al (new coefficients, may be
   NULL), b (right-hand side, nowned*nt, may be NULL) and A (borrowed wb_mat, bs = nt, valid until the tracers or
   the mesh change; may be NULL). */
int wb_tracer_setup_linear(wb_ctx *ctx, double dt, const double *al_last, const double *x_last,
                           const double *al_last2, const double *x_last2, const double *x_boundary, double *al,
                           double *b, wb_mat **A);
/* the whole auxiliary step: setup_linear, aux_pre_solve, PCSetUp, KSPSolve (zero initial guess).
   x (out): new mass fractions, nowned*nt. */
int wb_tracer_solve(wb_ctx *ctx, const wb_ksp_opts *ksp, int pc_type, int pc_nblocks, double dt,
                    const double *al_last, const double *x_last, const double *al_last2, const double *x_last2,
                    const double *x_boundary, double *al, double *x, int *its, int *reason);

/* ---- instrumentation (PetscLogEvent equivalents, src/profiling.F90:42-65) -- */
/* accumulated device time (ms) and call count of a named phase:
   "fluid_props", "cell_balances", "cell_inflows", "jacobian", "pc_setup", "ksp_solve", "mat_mult",
   "pc_apply", "fluid_trans", "tracer_setup", "tracer_solve" */
int wb_timer_get(wb_ctx *ctx, const char *name, double *ms, int64_t *count);
int wb_timer_reset(wb_ctx *ctx);
/* phase timers synchronise the stream at every phase end; switch them off for throughput runs */
int wb_timers_enable(int on);
/* The GMRES solve with block-Jacobi / ILU(0) sub-domains runs as ONE persistent kernel (the whole KSPSolve: no
   per-operation launches to time).  Its CTA 0 accumulates the device time it spends in each phase of the Krylov
   iterations; ns7 receives nanoseconds since the last reset: [0] SpMV + PC apply, [1] Gram-Schmidt dots, [2] grid /
   NVLink reduction of the dots, [3] multi-AXPY + norm + halo push, [4] reduction of the norm + Hessenberg update,
   [5] everything else (cycle starts, solution update), and [6] the number of iterations counted. */
int wb_ksp_fused_profile(wb_ctx *ctx, double *ns7, int reset);
/* number of kernels launched by this context since creation */
int64_t wb_launch_count(const wb_ctx *ctx);
/* stream the context launches on (cudaStream_t) */
void *wb_stream(wb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
