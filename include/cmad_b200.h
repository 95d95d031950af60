/*
 * cmad_b200 - B200-native (sm_100a) constitutive-update hot path for CMAD.
 *
 * C-ABI drop-in boundary.  Plain pointers and sizes only; no torch / JAX types.
 * Every entry point is asynchronous on the CUDA stream it is given (device
 * entry points) or self-contained (host-buffer entry points), returns 0 on
 * success or a CMADX_E* code, never throws, and allocates nothing persistent
 * except the opaque handles it returns.  The caller owns all buffers.
 *
 * The reference (sandialabs/cmad) is pure Python on JAX and has no FFI of its
 * own; each entry point below cites the reference *Python* interface it
 * replaces (file:line relative to the reference tree).  INTEGRATION.md shows
 * the jax.ffi / ctypes binding a maintainer adds on the reference side.
 *
 * Array layout ("SoA"): every per-point quantity is component-major, component
 * c of point i at  base[c*ld + i]  (ld >= n, in elements).  Symmetric tensors
 * use CMAD's packing order xx,xy,xz,yy,yz,zz (cmad/models/var_types.py:43-47,
 * 73-77); off-diagonal entries are *tensor* components (not engineering).
 */
#ifndef CMAD_B200_H
#define CMAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMADX_VERSION 100

/* ---- error codes -------------------------------------------------------- */
enum {
    CMADX_OK = 0,
    CMADX_EINVAL = 1,       /* bad enum / size / null pointer                 */
    CMADX_EUNSUPPORTED = 2, /* valid request the kernels do not implement     */
    CMADX_ECUDA = 3,        /* CUDA runtime error (see cmadx_last_cuda_error) */
    CMADX_ENOMEM = 4
};

/* ---- models: cmad/models/small_elastic_plastic.py:95, cmad/models/elastic.py:29 */
/* CMADX_MODEL_SMALL_RATE_ELASTIC_PLASTIC (cmad/models/small_rate_elastic_plastic.py): state =
 * [cauchy(6), alpha]; the `strain` rows of a batch carry the strain INCREMENT eps - eps_prev
 * (the reference forms it from U and U_prev); rotated material axes included.  FULL_3D: K1, the
 * forward history, K2 adjoint / direct, element blocks of any rule in the displacement and the
 * mixed u-p form (fe_rate.cu).  PLANE_STRESS / UNIAXIAL_STRESS (cmadx_mp_update only): n_xi = 8 /
 * 12 = [cauchy(6), alpha, stretches (1 | 2), off-axis delta strains (uniaxial: 3)].
 * Not carried: histories / objectives of the def-types, the Hessian path, K6. */
enum { CMADX_MODEL_SMALL_ELASTIC_PLASTIC = 0, CMADX_MODEL_ELASTIC = 1,
       CMADX_MODEL_SMALL_RATE_ELASTIC_PLASTIC = 2 };
/* ---- effective stress: cmad/models/effective_stress.py:16-27 */
/* CMADX_YIELD_BARLAT: Yld2004-18p, effective_stress.py:55-84 -> verification/functions.py:71-154
 * (the eigenvalues of two linear images of the stress); carried by the one-pass material-point
 * kernels (K1, forward history; the three def-types; the rate model), the calibration objectives
 * (K2 adjoint / direct) and the any-rule element kernels (K3 / K4 of SmallElasticPlastic); the
 * other entry points (Hessian, K6, rate-model element blocks) return CMADX_EUNSUPPORTED for it. */
enum { CMADX_YIELD_J2 = 0, CMADX_YIELD_HILL = 1, CMADX_YIELD_HOSFORD = 2, CMADX_YIELD_BARLAT = 3 };
/* ---- which two elastic constants are given (cmad/models/elastic_constants.py:54-104),
 *      values in sorted-key order "E" < "kappa" < "lambda" < "mu" < "nu"       */
enum {
    CMADX_EL_E_NU = 0, CMADX_EL_E_MU, CMADX_EL_E_KAPPA, CMADX_EL_E_LAMBDA,
    CMADX_EL_KAPPA_MU, CMADX_EL_KAPPA_NU, CMADX_EL_KAPPA_LAMBDA,
    CMADX_EL_LAMBDA_MU, CMADX_EL_LAMBDA_NU, CMADX_EL_MU_NU
};
/* ---- hardening terms present (cmad/models/hardening.py:25-34), bit mask */
enum { CMADX_HARD_VOCE = 1, CMADX_HARD_LINEAR = 2 };
/* ---- canonical parameter ids: columns of dC/dp are requested by id, in the
 *      order the caller wants them (the reference's order is the sorted-key
 *      pytree flatten order restricted to active leaves,
 *      cmad/parameters/parameters.py:214-243, 368-377)                        */
enum {
    CMADX_P_EL0 = 0, CMADX_P_EL1, CMADX_P_Y, CMADX_P_VOCE_S, CMADX_P_VOCE_D,
    CMADX_P_LIN_K, CMADX_P_HILL_F, CMADX_P_HILL_G, CMADX_P_HILL_H, CMADX_P_HILL_L,
    CMADX_P_HILL_M, CMADX_P_HILL_N, CMADX_P_HOSFORD_A, CMADX_P_Q00,
    /* Yld2004-18p: sp_12 sp_13 sp_21 sp_23 sp_31 sp_32 sp_44 sp_55 sp_66, the same nine of dp_*,
     * then the exponent (the caller orders columns; the reference's sorted-key order is
     * a, dp_*, sp_*) */
    CMADX_P_BARLAT_C0 = CMADX_P_Q00 + 9,
    CMADX_P_BARLAT_A = CMADX_P_BARLAT_C0 + 18,
    CMADX_NUM_PARAM_IDS = CMADX_P_BARLAT_A + 1
};
#define CMADX_MAX_ACTIVE 16

/* deformation types of the material-point path (cmad/models/deformation_types.py) */
#define CMADX_DEF_FULL_3D 0
#define CMADX_DEF_PLANE_STRESS 1
#define CMADX_DEF_UNIAXIAL_STRESS 2

/* ---- local Newton flavour ------------------------------------------------ */
enum {
    CMADX_NEWTON_TRACED = 0,     /* make_newton_solve,  cmad/models/nonlinear_solver.py:88-174 */
    CMADX_NEWTON_IMPERATIVE = 1  /* newton_solve(model), cmad/models/nonlinear_solver.py:14-85 */
};

/* cmadx_newton_t.flags: force the generic 7x7 Newton kernel even where the J2
 * radial-return specialisation applies (testing / A-B comparison) */
/* flags of cmadx_newton_t: bit 0 = always use the generic Newton kernels (no J2 radial-return
 * specialisation); bits 8..15 = `defer_after` K of the generic kernels' two-pass scheme: a point
 * that needs more than K Newton updates is re-solved by a second launch made of such points only
 * (warp-divergence control; same iterates, counts and flags - derivative outputs may differ at
 * rounding level between the two launches).  0 = library default (K = 2 for near-Tresca Hosford
 * exponents, off otherwise), 255 = off. */
enum {
    CMADX_NEWTON_F_GENERIC = 1,
    /* reduced Hosford batches: the plain one-thread-per-point kernel (128-thread blocks, warps
     * unsynchronised) instead of the lock-step blocks - A/B comparison; results are identical */
    CMADX_NEWTON_F_ONE_PASS = 2,
    /* material-point batches: use the streaming kernel with lane refill (mp_update_stream.cu)
     * instead of the one-pass generic kernels (one thread = one point, a warp waits for its
     * slowest lane, optional two-pass deferral).  Results are identical; measured on B200 it is
     * slower (profiles/r2e_k1_ab.jsonl) because the Jacobian / LU still runs partially filled
     * and every point pays one more residual evaluation in the drain - kept for A/B work */
    CMADX_NEWTON_F_STREAM = 4,
    /* material-point batches: generic kernel with warp-level parking (mp_update_queue.cu): lanes
     * that still iterate when most of their warp is done are parked in shared memory and resumed
     * in full warps; results equal the one-pass kernels' bit for bit */
    CMADX_NEWTON_F_QUEUE = 8,
    /* material-point batches: generic kernel with block-level hand-off (mp_update_cta.cu): a block
     * owns 512 points, the points that still iterate after `defer_after` updates (0: all plastic
     * points) are resumed by full warps from records in shared memory, every store is a full row */
    CMADX_NEWTON_F_CTA = 16
};
#define CMADX_NEWTON_DEFER_SHIFT 8
#define CMADX_NEWTON_DEFER_MASK 0xff00

/* Material = the reference's parameter pytree for one element block
 * (cmad/parameters/parameters.py:205-272), flattened to a POD. */
typedef struct cmadx_material {
    int32_t model;          /* CMADX_MODEL_*                                   */
    int32_t yield;          /* CMADX_YIELD_*                                   */
    int32_t elastic_pair;   /* CMADX_EL_*                                      */
    int32_t hardening_mask; /* CMADX_HARD_* bits                               */
    double elastic[2];      /* the two given constants, sorted-key order       */
    double Y;               /* plastic/flow stress/initial yield/Y             */
    double voce_S, voce_D;  /* plastic/flow stress/hardening/voce              */
    double linear_K;        /* plastic/flow stress/hardening/linear            */
    double hill[6];         /* F G H L M N                                     */
    double hosford_a;
    double Q[9];            /* "rotation matrix", row-major                    */
    double yield_tol;       /* cmad/models/small_elastic_plastic.py:116 (1e-14)*/
    double barlat[18];      /* sp_12 sp_13 sp_21 sp_23 sp_31 sp_32 sp_44 sp_55 sp_66, dp_* alike */
    double barlat_a;        /* plastic/effective stress/barlat/a                */
} cmadx_material_t;

/* Local Newton + line-search settings: make_newton_solve kwargs
 * (cmad/models/nonlinear_solver.py:88-100) and DEFAULT_LINE_SEARCH_SETTINGS
 * (cmad/util/line_search.py:40-46). */
typedef struct cmadx_newton {
    int32_t mode;           /* CMADX_NEWTON_*                                  */
    int32_t max_iters;
    int32_t ls_max_evals;   /* traced: probes of the quadratic line search, >= 1;
                               imperative: max_ls_evals of newton_solve's legacy line
                               search (cmad/models/nonlinear_solver.py:55-81), 0 = none
                               (the reference's default; cmadx_mp_update only)        */
    int32_t flags;          /* CMADX_NEWTON_F_* bits                           */
    double abs_tol, rel_tol;
    double ls_c1, ls_bmin, ls_bmax;
} cmadx_newton_t;

/* Per-point buffers of one batched update.  Inputs must be non-null; any
 * output may be null (skipped).  n_xi = 7 (small_elastic_plastic, FULL_3D:
 * plastic strain(6) + alpha) or 6 (elastic: cauchy(6)).                      */
typedef struct cmadx_mp_buffers {
    int64_t n;              /* points                                          */
    int64_t ld;             /* leading dimension (elements), >= n              */
    int32_t strain_comps;   /* FULL_3D: 6 = symmetric strain (grad_u := strain),
                               9 = grad_u row-major [k*3+j] = du_k/dx_j;
                               PLANE_STRESS: 3 = (e_xx, e_xy, e_yy), 4 = 2x2 grad_u row-major;
                               UNIAXIAL_STRESS: 1 = axial strain                 */
    int32_t def_type;       /* CMADX_DEF_* (cmad/models/deformation_types.py): 0 FULL_3D
                               (n_xi 7), PLANE_STRESS (n_xi 8: + out-of-plane stretch),
                               UNIAXIAL_STRESS (n_xi 9: + two off-axis stretches; 12 for the
                               rate model: + three off-axis delta strains).  For the
                               latter two the derivative outputs are w.r.t. the prescribed
                               symmetric components only: dsig_deps [6*ns], dxi_deps
                               [n_xi*ns] with ns = 3 (xx, xy, yy) or 1                  */
    const double* xi_prev;  /* [n_xi][ld]                                      */
    const double* strain;   /* [strain_comps][ld]                              */
    const double* xi_init;  /* [n_xi][ld] or NULL: starting iterate (NULL =>
                               xi_prev, the reference's behaviour after
                               advance_xi / in make_newton_solve).  With
                               max_iters = 0 every output is evaluated AT
                               (xi_init, xi_prev): Model.evaluate() semantics,
                               cmad/models/model.py:168-193                    */
    double* xi;             /* [n_xi][ld]   converged local state              */
    double* sigma;          /* [6][ld]      global cauchy stress               */
    double* dsig_deps;      /* [36][ld]     consistent tangent, (a*6+b), wrt the
                                           symmetric strain component b        */
    double* dxi_deps;       /* [n_xi*6][ld] IFT dxi/deps, (r*6+b)              */
    double* dC_dp;          /* [n_xi*n_active][ld]  dC/dp at (xi, xi_prev), (r*n_active+c) */
    double* dC_dxi;         /* [n_xi*n_xi][ld]                                 */
    double* dC_dxi_prev;    /* [n_xi*n_xi][ld]                                 */
    int32_t* iters;         /* [n] Newton updates taken (ii)                   */
    int32_t* flags;         /* [n] bit0: plastic branch at x0, bit1: at x*     */
    double* cnorm;          /* [n] final ||C||_2                               */
    double* C;              /* [n_xi][ld]   residual at the returned xi        */
    const double* strain_prev; /* [strain_comps][ld] or NULL.  small_rate_elastic_plastic only
                               (its residual sees eps(U) - eps(U_prev),
                               cmad/models/small_rate_elastic_plastic.py:41-51): NULL => the
                               `strain` rows already carry the increment; non-NULL => `strain`
                               and `strain_prev` are the total strains of this and the previous
                               step and the increment is formed on the device        */
} cmadx_mp_buffers_t;

int cmadx_version(void);
/* sizeof(cmadx_material_t), sizeof(cmadx_newton_t), sizeof(cmadx_mp_buffers_t),
 * sizeof(cmadx_mp_history_t), sizeof(cmadx_fe_block_t): lets a foreign-language
 * binding verify its struct mirrors. */
int cmadx_struct_sizes(int64_t* out5);
const char* cmadx_error_string(int code);
/* text of the last CUDA error seen by this thread ("" if none) */
const char* cmadx_last_cuda_error(void);

/* lambda, mu and d(lambda,mu)/d(elastic[0..1]) for a material:
 * out6 = {lambda, mu, dlam/de0, dlam/de1, dmu/de0, dmu/de1}.
 * Replaces ElasticConstants.from_params (cmad/models/elastic_constants.py:54-104). */
int cmadx_lame(const cmadx_material_t* mat, double* out6);

/* K1: batched constitutive update on DEVICE buffers, asynchronous on `stream`
 * (a cudaStream_t).  One call = for every point: local Newton solve of
 * C(xi; xi_prev, p, U) = 0, cauchy stress, consistent tangent, dC/dp, dC/dxi,
 * dC/dxi_prev.  Replaces, for a batch of points,
 *   make_newton_solve(model._residual)(xi_prev, params, U, U_prev) and its IFT
 *   rule (cmad/models/nonlinear_solver.py:88-174) or newton_solve(model)
 *   (:14-85); Model.cauchy / dC_dxi / dC_dxi_prev / dC_dp
 *   (cmad/models/model.py:121-166, 316-350) with active-column selection
 *   (cmad/parameters/parameters.py:368-377);
 * i.e. the bodies of the per-point loops cmad/cli/primal.py:158-175 and
 * cmad/objectives/mp_objective.py:73-87.                                      */
int cmadx_mp_update(const cmadx_material_t* mat, const cmadx_newton_t* newton,
                    const int32_t* active_pid, int32_t n_active,
                    const cmadx_mp_buffers_t* dev, void* stream);

/* Same operation on HOST buffers (pageable or pinned): copies inputs to the
 * device in chunks, runs K1, copies every non-null output back, overlapping
 * H2D / kernel / D2H on three streams.  Blocking.  `device` = CUDA ordinal.
 * `chunk_points` <= 0 picks a default.  Scratch is cached per device in a
 * handle created on first use and freed by cmadx_release_host_scratch().      */
int cmadx_mp_update_host(const cmadx_material_t* mat, const cmadx_newton_t* newton,
                         const int32_t* active_pid, int32_t n_active,
                         const cmadx_mp_buffers_t* host, int device,
                         int64_t chunk_points);
int cmadx_release_host_scratch(void);

/* ---- K2: material-point calibration objective over whole load histories ------
 * Buffers of a batch of independent material points ("experiments"), each with
 * its own strain history and calibration data.  Histories are stored as N+1
 * slabs (slab t = load step t, slab 0 = initial state / unused), every slab
 * component-major [comps][ld].  Replaces the per-point Python loops of
 * MPAdjointObjective / MPDirectObjective (cmad/objectives/mp_objective.py:
 * 62-147, 158-215) with the Calibration QoI J = sum_t 1/2 ||w o (cauchy - data)||^2
 * (cmad/qois/calibration.py:56-66), summed over all points of the batch.       */
typedef struct cmadx_mp_history {
    int64_t n;              /* points                                          */
    int64_t ld;             /* leading dimension of every slab row (>= n)      */
    int32_t nsteps;         /* N load steps                                    */
    int32_t strain_comps;   /* as in cmadx_mp_buffers_t; it also selects the deformation type:
                               6 | 9 FULL_3D (n_xi 7), 3 | 4 PLANE_STRESS (n_xi 8),
                               1 UNIAXIAL_STRESS (n_xi 9); xi_hist slabs have n_xi rows,
                               slab 0 holds the initial state (stretches = 1)        */
    const double* strain;   /* [N+1][strain_comps][ld]                         */
    const double* data;     /* [N+1][9][ld] cauchy data, 3x3 row-major         */
    double weight[9];       /* Calibration weight, 3x3 row-major, time-constant*/
    double* xi_hist;        /* [N+1][n_xi][ld]: slab 0 = initial state (input),
                               slabs 1..N written by cmadx_mp_forward_history  */
    int32_t* iters_hist;    /* [N+1][ld] or NULL: Newton iterations per step   */
    double* result;         /* [1 + n_active] device: J, dJ/dp (native params, before
                               Parameters.transform_grad)                      */
    double* workspace;      /* device scratch, cmadx_mp_objective_workspace_bytes() */
    double* J_point;        /* [n] or NULL: per-point objective                */
    int32_t qoi_kind;       /* CMADX_QOI_CALIBRATION (0, the default) or
                               CMADX_QOI_UNIAXIAL_CALIBRATION: UniaxialCalibration
                               (cmad/qois/uniaxial_calibration.py:69-85), UNIAXIAL_STRESS
                               histories only (strain_comps 1, uniaxial_stress_idx 0,
                               stretch_var_idx 2): J = sum_t 1/2 ||w_t o ([sigma_axial,
                               lambda_2 - 1, lambda_3 - 1] - data_t)||^2; rows 0..2 of
                               every data slab hold (sigma, e_2, e_3), `weight` is unused */
    int32_t reserved_;
    const double* weight_steps; /* UNIAXIAL_CALIBRATION: [N+1][3] device, the QoI's
                               per-step weights (weight[:, step]); same for all points */
} cmadx_mp_history_t;
enum { CMADX_QOI_CALIBRATION = 0, CMADX_QOI_UNIAXIAL_CALIBRATION = 1 };

/* bytes of `workspace` needed for n points and n_active parameters */
int64_t cmadx_mp_objective_workspace_bytes(int64_t n, int32_t n_active);

/* Forward pass: N successive K1 updates writing the converged state of step t
 * into xi_hist slab t (run_primal_pass / _forward_pass_with_storage,
 * cmad/cli/primal.py:129-176, cmad/objectives/mp_objective.py:62-89).          */
int cmadx_mp_forward_history(const cmadx_material_t* mat, const cmadx_newton_t* newton,
                             const cmadx_mp_history_t* hist, void* stream);

/* Adjoint gradient (reverse-time recurrence, mp_objective.py:112-142) and direct
 * (forward sensitivity, :174-210) gradient of the summed objective; both need
 * xi_hist filled by the forward pass and write `result` (deterministic
 * reduction, bit-reproducible run to run).                                    */
int cmadx_mp_objective_adjoint(const cmadx_material_t* mat, const int32_t* active_pid,
                               int32_t n_active, const cmadx_mp_history_t* hist, void* stream);
int cmadx_mp_objective_direct(const cmadx_material_t* mat, const int32_t* active_pid,
                              int32_t n_active, const cmadx_mp_history_t* hist, void* stream);

/* The same objective on HOST buffers (pageable or pinned): `host->strain` / `host->data` are host
 * histories in the slab layout above, `host->result` a host array of 1 + n_active doubles,
 * `host->J_point` an optional host array; xi_hist / iters_hist / workspace are ignored (device
 * scratch is cached per device as for cmadx_mp_update_host).  Chunks of points are pipelined
 * H2D -> forward history -> K2 -> D2H of 8 (1 + n_active) bytes; the chunk sums are added on the
 * host in chunk order.  Blocking.  adjoint != 0: reverse-time recurrence, else forward
 * sensitivities.  This is MPAdjointObjective.evaluate / MPDirectObjective.evaluate
 * (cmad/objectives/mp_objective.py:53-57) for a batch of experiments held in host memory.     */
int cmadx_mp_objective_host(const cmadx_material_t* mat, const cmadx_newton_t* newton,
                            const int32_t* active_pid, int32_t n_active,
                            const cmadx_mp_history_t* host, int adjoint, int device,
                            int64_t chunk_points);

/* Second-order pass: MPDirectAdjointObjective (cmad/objectives/mp_objective.py:218-343) -
 * J, dJ/dp and the Hessian d2J/dp2 of the summed objective in NATIVE parameter values
 * (before Parameters.transform_grad / transform_hessian, cmad/parameters/parameters.py:
 * 326-357).  Runs the adjoint pass (keeping phi_t), then the forward direct-adjoint pass in
 * which every Hessian entry is one hyper-dual evaluation of J_t + phi_t . C_t along
 * (dz/dp_i, dz/dp_j) - the templated forward-mode counterpart of Model.evaluate_hessians /
 * QoI.evaluate_hessians (cmad/models/model.py:134-148, 245-270).  `result` holds
 * 1 + n_active + n_active^2 doubles: J, grad, H row-major (symmetric).  `workspace` needs
 * cmadx_mp_hessian_workspace_bytes().  FULL_3D, PLANE_STRESS and UNIAXIAL_STRESS (by
 * strain_comps, as for the gradient entry points); deterministic.  SmallElasticPlastic: rotated
 * material axes in FULL_3D.  SmallRateElasticPlastic (small_rate_elastic_plastic.py:250-346): the
 * three def-types, identity and rotated axes; its QoI reads the state's stress, so both flag
 * settings give the same Hessian.
 * flags: 0 = the complete Hessian (equals the derivative of the gradient).
 * CMADX_HESS_F_REFERENCE_QOI_CROSS reproduces the reference entry for entry: its QoI builds
 * the mixed block d2J/dxi dparams by differentiating w.r.t. xi_PREV (cmad/qois/qoi.py:53-55),
 * which is zero for the Calibration QoI, so the terms d2J_dp_dxi . dxi_dp of
 * mp_objective.py:320,322 are missing there.  Both agree whenever no active parameter enters
 * the Cauchy stress (flow-stress parameters only - the reference's own tests); they differ
 * when elastic parameters are active.                                                     */
enum { CMADX_HESS_F_REFERENCE_QOI_CROSS = 1 };
int64_t cmadx_mp_hessian_workspace_bytes(int64_t n, int64_t ld, int32_t nsteps, int32_t n_active);
int cmadx_mp_objective_hessian(const cmadx_material_t* mat, const int32_t* active_pid,
                               int32_t n_active, const cmadx_mp_history_t* hist, int32_t flags,
                               void* stream);

/* ---- K3/K4: FE element-block kernels (displacement formulation, COUPLED mode) ----
 * One call = one mesh element block: for every element gather U, for every
 * integration point interpolate grad_u = U_e^T grad_N
 * (cmad/global_residuals/interpolation.py:49-55), run the local Newton from
 * xi_prev (cmad/global_residuals/global_residual.py:361-395), evaluate the
 * weak-form residual R = (grad_N @ sigma) w dv
 * (cmad/global_residuals/small_disp_equilibrium.py:112-118) and its IFT-corrected
 * tangent dR/dU, accumulate over the element's integration points
 * (per_element_R_and_K_coupled, cmad/fem/assembly.py:416-535) and emit the
 * per-element results in the layouts of assemble_element_block
 * (cmad/fem/assembly.py:616-732).  All arrays use the REFERENCE's layouts
 * (element-major, row-major trailing axes):
 *   geometry cache  cmad/fem/precompute.py:51-72,107-122
 *   index arrays    cmad/fem/kernel_arrays.py:58-114 (u_gather_eq == r_scatter_eq for
 *                   the single-field displacement formulation; dof = basis*3 + comp,
 *                   cmad/fem/assembly.py:142-165)
 *   xi history      flat-trailing (n_elems, n_ip, n_xi), cmad/fem/fe_problem.py:310-313
 * Element families: tet4 and hex8 with any volume rule (n_ip 1..64).  The reference's
 * defaults tet4 x 1 IP and hex8 x 8 IPs (cmad/fem/fe_problem.py:35-38) run the tuned
 * kernels; other rules (`discretization.quadrature.volume degree`, cmad/cli/common.py:
 * 497-540 - e.g. tet4 x 4, which the mixed formulation requires on tets, :379-391) run a
 * correctness-first kernel with one thread per element.  grad_N, K_elem and (tet4 x 1)
 * R_elem must be 32-byte aligned (256-bit vector loads/stores).                */
typedef struct cmadx_fe_block {
    int64_t n_elems;
    int64_t n_dofs;          /* length of U and R_global                         */
    int32_t n_basis;         /* 4 (tet4) | 8 (hex8)                              */
    int32_t n_ip;            /* 1 | 8                                            */
    const int32_t* elem_eq;  /* [n_elems][n_basis*3] global equation of (basis, comp) */
    const double* U;         /* [n_dofs] global displacement vector              */
    const double* xi_prev;   /* [n_elems][n_ip][n_xi]                            */
    const double* grad_N;    /* [n_elems][n_ip][n_basis][3] physical-frame       */
    const double* det;       /* [n_elems][n_ip] iso_jac_det (signed)             */
    const double* quad_w;    /* [n_ip]                                           */
    double* xi;              /* [n_elems][n_ip][n_xi] converged local state      */
    double* R_elem;          /* [n_elems][n_basis*3] or NULL                     */
    double* K_elem;          /* [n_elems][n_basis*3][n_basis*3] = the COO `vals`
                                stream of assemble_element_block, or NULL (K4:
                                residual-only, per_element_R_coupled :538-613)   */
    double* R_global;        /* [n_dofs] or NULL: atomic scatter-add of R_elem
                                (fast, not bit-reproducible; use
                                cmadx_segment_sum on R_elem for a deterministic R) */
    double* sigma;           /* [n_elems][n_ip][6] global cauchy at the IPs, or NULL
                                (evaluate_cauchy_at_ips, cmad/fem/postprocess.py:35-185) */
    int32_t* iters;          /* [n_elems][n_ip] or NULL                          */
    int32_t* flags;          /* [n_elems][n_ip] or NULL (bit0 entry, bit1 exit)  */
    const double* U_prev;    /* [n_dofs] displacement vector of the previous step: required by
                                small_rate_elastic_plastic blocks (their residual sees
                                eps(U) - eps(U_prev)), ignored (may be NULL) otherwise */
} cmadx_fe_block_t;

int cmadx_fe_block_assemble(const cmadx_material_t* mat, const cmadx_newton_t* newton,
                            const cmadx_fe_block_t* blk, void* stream);

/* ---- Mixed u-p (stabilised equal-order) formulation of SmallDispEquilibrium ------------
 * cmad/global_residuals/small_disp_equilibrium.py:87-111 (`mixed: true`, examples/
 * mixed_plastic.yaml): two residual blocks, u ("equilibrium") and p ("pressure"), block-major
 * dofs in ONE global vector U (cmad/fem/dof.py block_offsets).  The momentum stress is
 * dev(cauchy) - p I; R_p = (-(p + hydro)/kappa N - tau gradN.grad p) w dv with
 * hydro = kappa tr(eps), tau = stab_mult h^2 / (2 mu), h = RMS edge length of the element
 * (cmad/fem/mesh.py:624-636).  `blk` as for cmadx_fe_block_assemble: elem_eq / R_elem / K_elem
 * describe the u block (K_elem = the (u,u) COO stream); `mix` adds the pressure block and the
 * (u,p), (p,u), (p,p) streams, in the reference's (r, s) emit order
 * (cmad/fem/assembly.py:722-732).  blk->R_global / mix->R_global may alias the same [n_dofs]
 * vector (atomic scatter-add of both residual blocks).                                  */
typedef struct cmadx_fe_mixed {
    const int32_t* elem_eq_p; /* [n_elems][n_basis] global equation of the pressure dof  */
    const double* N;          /* [n_ip][n_basis] shape values at the IPs (shared)        */
    const double* h;          /* [n_elems] element size                                  */
    double stab_mult;         /* `stabilization multiplier`                              */
    double* R_p_elem;         /* [n_elems][n_basis] or NULL                              */
    double* K_up;             /* [n_elems][n_basis*3][n_basis] or NULL                   */
    double* K_pu;             /* [n_elems][n_basis][n_basis*3] or NULL                   */
    double* K_pp;             /* [n_elems][n_basis][n_basis] or NULL                     */
    double* R_global;         /* [n_dofs] or NULL: atomic scatter-add of R_p             */
} cmadx_fe_mixed_t;

int cmadx_fe_block_assemble_mixed(const cmadx_material_t* mat, const cmadx_newton_t* newton,
                                  const cmadx_fe_block_t* blk, const cmadx_fe_mixed_t* mix,
                                  void* stream);

/* ---- K6: forward sensitivities (JVP) of the element block at a converged state --------
 * For a tangent direction (dp over the active parameters, dxi_prev per point) at FIXED U:
 *   dxi = -A^{-1} (dC/dp dp + dC/dxi_prev dxi_prev),  A = dC/dxi at (xi_state, xi_prev),
 *   dR_e = sum_ip gradN (dcauchy/dxi dxi + dcauchy/dp dp) w dv.
 * This is what jax.jvp of the assembled residual w.r.t. (params, xi_prev) pushes through
 * make_newton_solve's custom_jvp rule (cmad/models/nonlinear_solver.py:158-171) inside the
 * FE Newton's IFT rule (cmad/fem/nonlinear_solver.py:490-537), and the per-step body of a
 * direct FE sensitivity recurrence.  `blk` as in cmadx_fe_block_assemble with
 * K_elem == NULL; blk->xi receives dxi, blk->R_elem / R_global receive dR.
 * xi_state: the converged local state [n_elems][n_ip][7] (the `xi` output of the primal
 * call); dxi_prev: [n_elems][n_ip][7] or NULL (zero); dp_host: n_active doubles (host).
 * dU_global ([n_dofs] or NULL) adds a displacement direction: dC/deps deps enters dxi
 * and dcauchy/deps deps enters dR (then dR includes K dU) - the "jvp of the xi output"
 * of cmad/fem/nonlinear_solver.py:534-538 once the global sensitivity dU is known.     */
int cmadx_fe_block_jvp(const cmadx_material_t* mat, const int32_t* active_pid, int32_t n_active,
                       const double* dp_host, const cmadx_fe_block_t* blk,
                       const double* xi_state, const double* dxi_prev, const double* dU_global,
                       void* stream);

/* Reverse mode (VJP), the transpose of cmadx_fe_block_jvp: given the cotangent Rbar of the
 * assembled residual (a nodal adjoint vector, [n_dofs]) and the cotangent xibar of the
 * converged local state ([n_elems][n_ip][7] or NULL), computes
 *   pbar[c] = sum over points of (dC/dp)^T mu + (dcauchy/dp)^T sbar   (device, n_active),
 *   blk->xi := xibar_prev = (dC/dxi_prev)^T mu,     mu = -A^{-T}(xibar + (dcauchy/dxi)^T sbar),
 * i.e. one step of a discrete FE adjoint (what jax.grad obtains by transposing the JVP
 * above, cmad/cli/gradient.py:74-82).  pbar is reduced in fixed order (bit-reproducible);
 * `workspace` needs cmadx_fe_vjp_workspace_bytes().  K_elem/R_elem/R_global are ignored. */
int64_t cmadx_fe_vjp_workspace_bytes(int64_t n_elems, int32_t n_ip, int32_t n_active);
int cmadx_fe_block_vjp(const cmadx_material_t* mat, const int32_t* active_pid, int32_t n_active,
                       const cmadx_fe_block_t* blk, const double* xi_state,
                       const double* Rbar_global, const double* xibar, double* pbar_dev,
                       double* workspace, void* stream);

/* K6 for the mixed u-p block (cmad/global_residuals/small_disp_equilibrium.py:87-111; the
 * sensitivities `cmad gradient` needs on examples/mixed_plastic.yaml-style decks): the same
 * tangent / cotangent maps over BOTH residual blocks.  `mix` as in
 * cmadx_fe_block_assemble_mixed with K_up = K_pu = K_pp = NULL.
 * JVP: blk->R_elem / R_global receive dR_u = sum_ip gradN (dev(d cauchy) - dp I) w dv and
 *      mix->R_p_elem / R_global receive
 *      dR_p[a] = sum_ip ((p dkappa/kappa^2 - dp/kappa - tr(d eps)) N_a
 *                        + (tau dmu/mu gradN_a.grad p - tau gradN_a.grad dp)) w dv;
 *      dU_global covers the block-major (u, p) dofs, (dp, d eps) are interpolated from it.
 * VJP: Rbar_global covers both blocks; the momentum cotangent is projected (dev is
 *      self-adjoint) before it enters the local adjoint solve, and the pressure rows add
 *      their (kappa, mu) cotangents to the elastic parameters' entries of pbar.          */
int cmadx_fe_block_jvp_mixed(const cmadx_material_t* mat, const int32_t* active_pid, int32_t n_active,
                             const double* dp_host, const cmadx_fe_block_t* blk,
                             const cmadx_fe_mixed_t* mix, const double* xi_state,
                             const double* dxi_prev, const double* dU_global, void* stream);
int cmadx_fe_block_vjp_mixed(const cmadx_material_t* mat, const int32_t* active_pid, int32_t n_active,
                             const cmadx_fe_block_t* blk, const cmadx_fe_mixed_t* mix,
                             const double* xi_state, const double* Rbar_global, const double* xibar,
                             double* pbar_dev, double* workspace, void* stream);

/* Displacement cotangent of the converged block (the piece a discrete FE adjoint through the
 * load steps needs besides the assembled tangent): per integration point p,
 *   Ubar_ip[p][(a,k)] = contribution of point p to (d xi/dU)^T xibar + (d R_u/dU |total)^T Rbar
 * ([n_elems*n_ip][n_basis*3]; sum the rows of an element - or scatter them with a
 * cmadx_segment_sum plan over the equation row repeated per point - to obtain the nodal
 * vector; bit-reproducible that way).  This transposes the displacement direction of
 * cmadx_fe_block_jvp: what jax.grad obtains from the `xi` output of the FE Newton's IFT rule
 * (cmad/fem/nonlinear_solver.py:534-538).  Rbar_global and xibar may each be NULL (zero);
 * blk->xi receives xibar_prev as in cmadx_fe_block_vjp.  `mix` (or NULL): for the mixed u-p
 * formulation the momentum cotangent is projected (dev); the pressure rows' dependence on U
 * (K_pu) is state-independent and is NOT included - it is in the assembled tangent.        */
int cmadx_fe_block_vjp_disp(const cmadx_material_t* mat, const cmadx_fe_block_t* blk,
                            const cmadx_fe_mixed_t* mix, const double* xi_state,
                            const double* Rbar_global, const double* xibar, double* Ubar_ip,
                            void* stream);

/* ---- Post-processing at a stored state: evaluate_cauchy_at_ips ------------------------
 * model.cauchy(xi, xi_prev, params, U_ip, U_ip_prev) at every (element, IP) of a COUPLED block
 * from the converged local state (cmad/fem/postprocess.py:35-185): sigma [n_elems][n_ip][6],
 * global axes, packed xx,xy,xz,yy,yz,zz.  Uses blk->elem_eq, U, grad_N and the counts only.  */
int cmadx_fe_cauchy_at_ips(const cmadx_material_t* mat, const cmadx_fe_block_t* blk,
                           const double* xi_state, double* sigma, void* stream);

/* ---- Embedded Dirichlet BCs on the deduplicated COO tangent ---------------------------
 * _embedded_bc_enforce + _embedded_residual (cmad/fem/sparse_solve.py:1058-1174) as two
 * HBM-bound passes over device arrays.  The plan is built once per (pattern, prescribed set):
 * rows/cols = the UNIQUE (deduplicated) pattern the COO dedup plan writes to.
 *   K_emb[e] = K[e] where both indices are free or e is a prescribed diagonal, else 0
 *              (the prescribed diagonal keeps the assembled K_ii, as the reference's);
 *   r[i]     = R[i] + sum_{j prescribed} K[i,j] (val_j - U_j)   on free rows (entry order),
 *   r[i]     = K_ii (U_i - val_i)                               on prescribed rows.
 * presc_vals: device, in the order of presc_idx.  K_emb may alias K; r must not alias R.   */
typedef struct cmadx_embedded_plan cmadx_embedded_plan_t;
int cmadx_embedded_plan_create(const int64_t* rows_host, const int64_t* cols_host, int64_t nnz,
                               int64_t n_dofs, const int64_t* presc_idx_host, int64_t n_presc,
                               cmadx_embedded_plan_t** plan);
int cmadx_embedded_plan_destroy(cmadx_embedded_plan_t* plan);
int cmadx_embedded_apply(const cmadx_embedded_plan_t* plan, const double* K_data, const double* R,
                         const double* U, const double* presc_vals, double* r_out,
                         double* K_emb_out, void* stream);

/* ---- K5: deterministic segment sums (R scatter-add, COO dedup) -----------------
 * out[s] = sum of vals[i] over all items i with seg_of_item[i] == s, summed in
 * increasing item order (bit-reproducible, and the order a sequential
 * `.at[idx].add(vals)` uses).  Replaces R_block.at[eq].add(R_flat)
 * (cmad/fem/assembly.py:715-720) with seg = r_scatter_eq, and the COO dedup
 * unique_data.at[coo_dedup_scatter].add(vals) (cmad/fem/assembly.py:906-909,
 * 1026-1070) with seg = coo_dedup_scatter.  The plan (a CSR of item lists per
 * segment) is built once per mesh on the host and lives on the current device. */
typedef struct cmadx_segment_plan cmadx_segment_plan_t;
int cmadx_segment_plan_create(const int64_t* seg_of_item_host, int64_t n_items,
                              int64_t n_segments, cmadx_segment_plan_t** plan);
int cmadx_segment_plan_destroy(cmadx_segment_plan_t* plan);
/* accumulate != 0: out[s] += sum (e.g. adding a block's R into the global R) */
int cmadx_segment_sum(const cmadx_segment_plan_t* plan, const double* vals_dev,
                      double* out_dev, int accumulate, void* stream);

/* Pack / unpack of a sorted dof subset (the interface dofs of an element partition) around the
 * halo all-reduce of the assembled residual: packed[i] = src[index[i]] / dst[index[i]] = packed[i].
 * `index` is a device array of n distinct positions.                                        */
int cmadx_index_gather(const int64_t* index_dev, int64_t n, const double* src_dev, double* packed_dev,
                       void* stream);
int cmadx_index_scatter(const int64_t* index_dev, int64_t n, const double* packed_dev, double* dst_dev,
                        void* stream);

/* Raw AD products of the reference's Model object at GIVEN states (Model.evaluate semantics,
 * cmad/models/model.py:121-160,168-193): dC/dU (jacfwd over U), dcauchy/dxi, dcauchy/dU and
 * dcauchy/dparams.  SmallElasticPlastic, FULL_3D, every effective stress, rotated axes included.
 * Component-major device arrays, leading dimension ld >= n; a NULL output is skipped.  U-derivatives
 * are with respect to the SYMMETRIC strain components (both entries moving; the reference's
 * single-entry columns are half of them off the diagonal).  dC/dU_prev and dcauchy/dxi_prev vanish
 * identically for this model and have no output.  dsig_dp carries the elastic constants (zero
 * columns for the flow-stress / yield-surface leaves, which cauchy does not see); rotation-matrix
 * leaves there: CMADX_EUNSUPPORTED.                                                           */
typedef struct cmadx_mp_partials {
    int64_t n, ld;
    int32_t strain_comps;   /* 6 = symmetric strain, 9 = grad_u row-major       */
    int32_t reserved;
    const double* xi;       /* [7][ld] state the products are evaluated at      */
    const double* xi_prev;  /* [7][ld]                                          */
    const double* strain;   /* [strain_comps][ld]                               */
    double* dC_deps;        /* [7*6][ld]  (r*6+b)                               */
    double* dsig_dxi;       /* [6*7][ld]  (a*7+c), global cauchy                */
    double* dsig_deps;      /* [36][ld]   (a*6+b), partial (xi held fixed)      */
    double* dsig_dp;        /* [6*n_active][ld] (a*n_active+c), partial         */
} cmadx_mp_partials_t;
int cmadx_mp_model_partials(const cmadx_material_t* mat, const int32_t* active_pid, int32_t n_active,
                            const cmadx_mp_partials_t* p, void* stream);

/* Batched eigen-decomposition of symmetric 3x3 tensors: replaces `sorted_eigen_decomposition`
 * (cmad/util/jax_eigen_decomposition.py:86-171).  Component-major device arrays with leading
 * dimension ld >= n: A6 rows xx,xy,xz,yy,yz,zz; w rows = eigenvalues in ascending order; V rows
 * 3 m + k = component m of the k-th eigenvector (NULL: eigenvalues only).                     */
int cmadx_sym3_eigh(int64_t n, int64_t ld, const double* A6_dev, double* w_dev, double* V_dev, void* stream);

/* debugging aid: how many points the last J2 radial-return launch on `stream`
 * handed back to the generic kernel (synchronises the stream); -1 if none ran */
int64_t cmadx_debug_bail_count(void* stream);

/* test / bench support: the kernel-side forms of a material and of the Newton settings (Lame
 * constants and their derivatives, integer Hosford exponent, line-search constants ...) as
 * opaque bytes, so that a host build of the per-point routines (the CPU baseline
 * oracle/j2_host.cpp) runs with exactly the constants the kernels get.  `sizes[2]` receives the
 * byte sizes; either buffer may be NULL to query them. */
int cmadx_debug_device_structs(const cmadx_material_t* mat, const cmadx_newton_t* newton,
                               void* dev_mat, void* dev_newton, int64_t* sizes);

/* number of kernel launches issued by this library since load (all threads) */
int64_t cmadx_launch_count(void);

/* FP64 FMA peak micro-benchmark (dependent-chain-free DFMA loop): returns
 * achieved TFLOP/s on the current device through *tflops.  Used by bench.py to
 * obtain the FP64 roofline denominator MEASURED_PEAKS.json lacks.             */
int cmadx_fp64_peak(int iters, double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMAD_B200_H */
