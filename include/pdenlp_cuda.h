/*
 * pdenlp_cuda.h — C ABI of libpdenlp_cuda.so
 *
 * B200-native (sm_100a, FP64) replacement for the derivative-evaluation hot
 * path of PDENLPModels.jl: the NLPModels callbacks of a GridapPDENLPModel.
 * Citations are relative to /root/reference (PDENLPModels.jl v0.3.5).
 *
 * The reference has no FFI for this path (it is pure Julia); every entry point
 * below names the Julia method whose *body* it replaces.  The Julia wrapper
 * (julia/PDENLPCuda.jl) keeps the method signatures, the
 * @lencheck DimensionError checks and the Counters bookkeeping on its side and
 * issues exactly one ccall per callback.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types cross the boundary.
 *   - every function returns an int status: PDENLP_OK (0) or a negative
 *     PDENLP_E* code; pdenlp_last_error() returns a thread-local message.
 *   - all floating point is IEEE FP64; all dof ids are 1-based, signed int32
 *     (negative id = Dirichlet dof, -id indexes the Dirichlet-value array —
 *     Gridap's PosNegReindex, src/PDENLPModels.jl:27-43).
 *   - x layout: x = [kappa (nparam) ; y free dofs (nfree_y) ; u free dofs (nfree_u)]
 *     (src/GridapPDENLPModel.jl:249-251, src/util_functions.jl:9-17).
 *   - vector arguments of the callbacks are DEVICE pointers unless the model was
 *     created with PDENLP_MEM_HOST, in which case they are HOST pointers and the
 *     library stages them (pinned bounce buffers, copies inside the call).
 *   - a model handle owns the cells [0, ncells) of the arrays it was created
 *     with.  Multi-GPU: one handle per GPU over a contiguous cell slab; the
 *     COO value streams of the slabs concatenate to the global streams in rank
 *     order, so jac_coord/hess_coord need no communication.
 *   - callbacks on one handle are not re-entrant (same as the reference, whose
 *     model owns its workspace: src/GridapPDENLPModel.jl:56-73).
 *   - there is NO CPU fallback: every entry point fails with PDENLP_ECUDA when
 *     no sm_100 device / driver is present.
 */
#ifndef PDENLP_CUDA_H
#define PDENLP_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDENLP_ABI_VERSION 3

/* status codes */
#define PDENLP_OK 0
#define PDENLP_EINVAL -1       /* bad argument / inconsistent description   */
#define PDENLP_EUNSUPPORTED -2 /* (integrand, element) outside the closed set */
#define PDENLP_ECUDA -3        /* CUDA runtime error (message has details)   */
#define PDENLP_ENOMEM -4

/* closed, named set of weak-form integrands (SURVEY §8a catalogue) */
enum pdenlp_integrand {
  /* f = 0.5(yd-y)^2 + 0.5*alpha*u^2 ; res = grad(v).grad(y) - v*u - v*h
     README.md:58-98, src/GridapPDENLPModel.jl:164-204.  params[0]=alpha */
  PDENLP_MEMBRANE = 1,
  /* f as above ; res = grad(v).grad(y) + sinh(y)*v - u*v - v*h
     docs/src/poisson-boltzman.md:30-67.  params[0]=alpha */
  PDENLP_POISSON_BOLTZMANN = 2,
  /* f as above ; res = -nu*grad(v).grad(y) + (grad(y).1)*v - v*u - v*h
     test/problems/burger1d.jl:36-46, src/util_functions.jl:109-110.
     params[0]=alpha, params[1]=nu */
  PDENLP_BURGER_LIN = 3,
  /* nparam=1: f = ... + 0.5*(k1-params[2])^2 ;
     res = nu*grad(v).grad(y) - k1 + v*y*(grad(y).1) - v*u - v*h
     test/problems/burger1d_param.jl:30-46.  params[0]=alpha, [1]=nu, [2]=0.08 */
  PDENLP_BURGER_NL = 4,
  /* nparam=2: f = 0.5(sol-y)^2 + 0.5(k1-1)^2 + 0.5(k2-1)^2 ;
     res = k1*grad(v).grad(y) - v*f*k2 [- v*u when a control space is given]
     test/problems/poissonmixed.jl:33-44 */
  PDENLP_KAPPA_POISSON = 5,
  /* unconstrained (no res, ncon = 0): f = 0.5*grad(y).grad(y) - w*y, w = coef source
     test/problems/penalizedpoisson.jl:31-38 */
  PDENLP_PENALIZED_POISSON = 6,
  /* unconstrained, two fields: f = 0.5*(ubis-u)^2 + 0.5*y^2, ubis = coef target
     test/problems/basicunconstrained.jl:21-25 */
  PDENLP_BASIC_UNCONSTRAINED = 7,
  /* ---- multi-field state / control spaces (MultiFieldFESpace); 1-D, dt(y,v) = (grad(y).1)*v ----
     state (I,S), no control: f = 0.5*I*I ;
     res = dt(I,p) + dt(S,q) - p*(a*S*I - b*I) - q*(mu - b*I - a*S*I)
     test/problems/controlsir.jl:20-33.  params = (a, b, mu) */
  PDENLP_CONTROL_SIR = 8,
  /* state (cf,pf), control u: f = kp*pf ;
     res = dt(cf,p) + dt(pf,q) - p*(kp*pf*(1-cf) - kr*cf*(1-cf-pf)) + q*(u*kr*cf*(1-cf-pf) - kp*pf*pf)
     test/problems/cellincrease.jl:15-40.  params = (kp, kr) */
  PDENLP_CELL_INCREASE = 9,
  /* state (I,S), controls (bf,cf): f = 0.5*(bf*w0 - w1)^2 + 0.5*(cf*w0 - w2)^2, w0 = coef target ;
     res = dt(I,p) + dt(S,q) - p*(bf*S*I - cf*I) + q*bf*S*I
     test/problems/dynamicsir.jl:36-58.  params = (w1, w2) */
  PDENLP_DYNAMIC_SIR = 10,
  /* unconstrained, state (phi,theta): f = a^2*phi'^2 + (c + a*cos(phi))^2 * theta'^2
     test/problems/torebrachistochrone.jl:25-28.  params = (a, c) */
  PDENLP_TORE_BRACHISTOCHRONE = 11,
  /* ---- NoFETerm objectives: obj = fk(kappa), no integral (src/additional_obj_terms.jl:59-106); the
     objective Hessian has no FE block, its kappa-block is the p x p lower triangle ----
     nparam=1: fk = 0.5*dot(k .- 1, k .- 1) ; res = k1*grad(v).grad(y) - v*f
     test/problems/poissonparam.jl:35-44 */
  PDENLP_POISSON_PARAM = 12,
  /* state (I,S), NoFETerm() = 0: res = dt(I,p) + dt(S,q) - p*(a*S*I - b*I) - q*(b*I - a*S*I)
     test/problems/sis.jl:21-34.  params = (a, b) */
  PDENLP_SIS = 13,
  /* nparam=2, plain MixedEnergyFETerm (inde = false: objective kappa-block with cross rows):
     f = k1*(sol-y)^2 + 0.5*(k2-1)^2 ; res = k1*grad(v).grad(y) - v*f*k2
     test/problems/poissonmixed2.jl:35-44 */
  PDENLP_KAPPA_POISSON2 = 14,
  /* TEST ONLY: PDENLP_POISSON_BOLTZMANN with its shortcut traits deliberately mis-declared; pdenlp_create must
     reject it with PDENLP_EINVAL (create-time self-check, see pdenlp_selfcheck_error) */
  PDENLP_SELFTEST_WRONG_TRAIT = 1000
};

/* coefficient fields sampled at quadrature points, coef[k][cell][q] */
#define PDENLP_COEF_TARGET 0 /* yd / sol */
#define PDENLP_COEF_SOURCE 1 /* h / f    */
#define PDENLP_MAX_PARAMS 8

/* where the callback vectors live */
#define PDENLP_MEM_DEVICE 0
#define PDENLP_MEM_HOST 1

/*
 * Flattened model description — what the Julia layer extracts from Gridap once,
 * right after the unchanged GridapPDENLPModel constructor
 * (src/additional_constructors.jl:130-266).  All array pointers are HOST
 * pointers; pdenlp_create copies them to the device.
 */
typedef struct pdenlp_desc {
  int32_t abi_version;   /* PDENLP_ABI_VERSION */
  int32_t integrand;     /* enum pdenlp_integrand */
  int32_t dim;           /* 1, 2, 3 */
  int32_t ny;            /* local dofs of ONE field of the state space Ypde (= test space Xpde) */
  int32_t nu;            /* local dofs of ONE field of the control space Ycon, 0 if VoidFESpace */
  int32_t nq;            /* quadrature points per cell */
  int32_t nparam;        /* number of algebraic unknowns kappa */
  int32_t memspace;      /* PDENLP_MEM_DEVICE | PDENLP_MEM_HOST */
  int32_t device;        /* CUDA device ordinal, -1 = current device */
  int32_t nfields_y;     /* fields of a MultiFieldFESpace Ypde; 0 or 1 = single field */

  int64_t ncells;        /* cells owned by this handle (a contiguous slab) */
  int64_t nfree_y;       /* num_free_dofs(Ypde) = ncon */
  int64_t nfree_u;       /* num_free_dofs(Ycon) */
  int64_t ndir_y;        /* Dirichlet dofs of Ypde */
  int64_t ndir_u;        /* Dirichlet dofs of Ycon */

  const int32_t* cell_dofs_y; /* [ncells][nfields_y*ny], get_cell_dof_ids(Ypde) */
  const int32_t* cell_dofs_u; /* [ncells][nfields_u*nu], get_cell_dof_ids(Ycon) (no multi-field offset) */
  const double* dir_y;        /* [ndir_y] Dirichlet values of the TRIAL space Ypde */
  const double* dir_u;        /* [ndir_u] */

  /* reference-FE tables at the quadrature points.
     Uniform geometry (cell_jinvT == NULL: every cell is the same affine image of the reference cell, e.g. a
     CartesianDiscreteModel without `map`): grad_y holds the PHYSICAL gradients (J^-T * reference gradient, the
     same for every cell) and wdet = quadrature weight * det(J).
     Per-cell geometry (cell_jinvT != NULL, any Gridap triangulation): grad_y holds the REFERENCE gradients and
     wdet the reference quadrature weights; the library applies each cell's J^-T and det(J). */
  const double* phi_y;   /* [nq][ny]      */
  const double* grad_y;  /* [nq][ny][dim] */
  const double* phi_u;   /* [nq][nu]      */
  const double* wdet;    /* [nq] */

  int32_t ncoef;         /* number of coefficient fields (<= 2) */
  int32_t nfields_u;     /* fields of a MultiFieldFESpace Ycon; 0 or 1 = single field */
  const double* coef;    /* [ncoef][ncells][nq] */
  double params[PDENLP_MAX_PARAMS];

  /* multi-field spaces only (nfields_y > 1 or nfields_u > 1), else may be NULL:
     [nfields_y + nfields_u] free / Dirichlet dof counts per field, state fields first.  All fields
     of a space share the reference element (ny / nu local dofs); cell_dofs_y is then
     [ncells][nfields_y*ny] field-major with each field's OWN ids (get_cell_dof_ids of the single
     field spaces), dir_y the per-field Dirichlet values concatenated, nfree_y / ndir_y the sums;
     likewise for the control space.  x = [kappa | y fields | u fields]
     (ConsecutiveMultiFieldStyle, src/additional_constructors.jl:236-242). */
  const int64_t* field_nfree;
  const int64_t* field_ndir;

  /* ---- per-cell geometry (ABI v3): "weights and Jacobians" of the cell maps, evaluated by Gridap at the
     quadrature points (get_cell_map(trian), src/additional_constructors.jl:130-175 accepts any triangulation).
     geom_nq = 1: affine cells, one Jacobian per cell;  geom_nq = nq: one per cell and quadrature point.
     cell_jinvT[c][g][d][e] = (J^-1)[e][d], so that  grad_x phi = cell_jinvT * grad_xi phi;  cell_detj[c][g] = |det J|.
     NULL (both) selects the uniform fast path above. */
  const double* cell_jinvT; /* [ncells][geom_nq][dim][dim] */
  const double* cell_detj;  /* [ncells][geom_nq]           */
  int32_t geom_nq;          /* 0 (no per-cell geometry), 1 or nq */
  int32_t flags;            /* PDENLP_FLAG_* (0 = defaults; the former reserved word) */
} pdenlp_desc;

/* Deterministic accumulation.  By default the dof-indexed outputs (grad!, cons!, jprod!, jtprod!, hprod!, the dense
   kappa blocks) are summed with floating-point atomics: equal to the reference's cell-ordered sums up to round-off, not
   bitwise reproducible from run to run.  With this flag the cells are coloured once at create time (greedy: no two
   cells of a colour share a free dof) and those callbacks run one launch per colour with plain read-modify-write
   stores, so every entry is summed in a fixed order — bitwise identical results on every run, at 1.5-3x the time of
   the atomic kernels.  jac_coord!/hess_coord! FE segments and obj are deterministic either way. */
#define PDENLP_FLAG_DETERMINISTIC 1
/* Templates off.  For affine problems on identical cells jac_coord! / hess_coord! copy the COO image of one template
   cell to every cell with the same free-dof pattern instead of evaluating it.  With this flag every cell
   is evaluated by the generic kernels: same values, entry by entry; used by the tests that compare both paths over whole
   full-size outputs, and available to callers who want the reference's cell-by-cell evaluation literally. */
#define PDENLP_FLAG_NO_TEMPLATES 2

typedef struct pdenlp_model pdenlp_model; /* opaque */

/* sizes derived at create time; mirrors NLPModelMeta/PDENLPMeta fields
   (src/additional_constructors.jl:225-261) */
typedef struct pdenlp_info {
  int64_t nvar;       /* nparam + nfree_y + nfree_u */
  int64_t ncon;       /* nfree_y */
  int64_t nnzj;       /* LOCAL (this slab) Jacobian entries: k-block + y-block + u-block */
  int64_t nnzj_k;     /* ncon*nparam (dense block, only on the handle with own_kblocks) */
  int64_t nnzj_y;     /* y-block entries of this slab */
  int64_t nnzj_u;     /* u-block entries of this slab */
  int64_t nnzh;       /* LOCAL Hessian entries: obj k + obj FE + cons k + cons FE */
  int64_t nnzh_obj;   /* obj k-block + obj FE block (pdemeta.nnzh_obj) */
  int64_t nnzh_k_obj; /* get_nnz_hess_k, src/hessian_struct_nnzh_functions.jl:61-68 */
  int64_t nnzh_fe;    /* entries of ONE FE block (the constraint block; the objective block is equal or absent) */
  int64_t nnzh_k_con; /* p(p+1)/2 + (n-p)p, src/GridapPDENLPModel.jl:362 */
  int64_t nnzh_fe_obj;/* entries of the OBJECTIVE FE block: nnzh_fe, or 0 for a NoFETerm objective
                         (src/hessian_struct_nnzh_functions.jl:41-43); nnzh_obj = nnzh_k_obj + nnzh_fe_obj,
                         nnzh = nnzh_obj + nnzh_k_con + nnzh_fe */
} pdenlp_info;

const char* pdenlp_last_error(void);
int pdenlp_abi_version(void);
/* build provenance: 16 hex digits of the sha256 over the library's sources and headers, baked in at compile time
   (pdenlpmodels.jl_b200/build.py: source_hash).  The host layer recomputes it from the tree and refuses a library
   built from other sources. */
const char* pdenlp_build_id(void);

/* replaces nothing in the reference: new flattening hook run once after the
   constructor. */
int pdenlp_create(const pdenlp_desc* desc, pdenlp_model** out);
int pdenlp_destroy(pdenlp_model* m);
int pdenlp_get_info(const pdenlp_model* m, pdenlp_info* info);

/* Create-time self-check.  Integrand traits select shortcuts (a constraint FE Hessian block that is structurally zero
   is zero-filled instead of computed; integrands whose FE-field derivatives are the same at every quadrature point of
   a cell use pre-integrated reference tensors).  pdenlp_create runs the shortcut and the generic per-point kernels
   on a sample of the model's own cells with pseudo-random x, lambda, v and fails with PDENLP_EINVAL if they differ by
   more than 1e-12 (relative); this returns the worst relative difference it saw (0 when no shortcut applies). */
int pdenlp_selfcheck_error(const pdenlp_model* m, double* max_rel_err);

/* run the callbacks on this CUDA stream (a cudaStream_t cast to void*);
   NULL = the legacy default stream */
int pdenlp_set_stream(pdenlp_model* m, void* cuda_stream);
int pdenlp_sync(pdenlp_model* m);

/* device memory for callers without a CUDA binding (Julia without CUDA.jl) */
int pdenlp_malloc(void** ptr, int64_t bytes);
int pdenlp_free(void* ptr);
int pdenlp_memcpy_h2d(void* dst, const void* src, int64_t bytes);
int pdenlp_memcpy_d2h(void* dst, const void* src, int64_t bytes);
/* page-locked host memory: host vectors handed to a PDENLP_MEM_HOST model should
   live here so the staging copies run at full PCIe speed */
int pdenlp_host_alloc(void** ptr, int64_t bytes);
int pdenlp_host_free(void* ptr);

/* obj(nlp, x)                          src/GridapPDENLPModel.jl:245-255
   f is a HOST pointer to one double (partial sum of this slab). */
int pdenlp_obj(pdenlp_model* m, const double* x, double* f);

/* grad!(nlp, x, g)                     src/GridapPDENLPModel.jl:257-266
   g[nvar] is zero-filled then accumulated (assemble_vector! semantics). */
int pdenlp_grad(pdenlp_model* m, const double* x, double* g);

/* objgrad!(nlp, x, g): NLPModels composes it from obj and grad! (the reference does not override it);
   here one sweep over the cells yields both.  f: HOST pointer to one double. */
int pdenlp_objgrad(pdenlp_model* m, const double* x, double* f, double* g);
/* objcons!(nlp, x, c): likewise the objective and the residual from one sweep */
int pdenlp_objcons(pdenlp_model* m, const double* x, double* f, double* c);

/* cons_nln!(nlp, x, c)                 src/GridapPDENLPModel.jl:433-447,
   _from_terms_to_residual!             src/jacobian_functions.jl:23-58 */
int pdenlp_cons(pdenlp_model* m, const double* x, double* c);

/* jac_nln_coord!(nlp, x, vals)         src/GridapPDENLPModel.jl:567-584,
   _jac_coord!                          src/jacobian_functions.jl:122-206
   vals[nnzj] in jac_nln_structure! order: [k-block | y-block | u-block]. */
int pdenlp_jac_coord(pdenlp_model* m, const double* x, double* vals);

/* hess_coord!(nlp, x, vals; obj_weight)     src/GridapPDENLPModel.jl:268-301
   hess_coord!(nlp, x, lambda, vals; obj_weight)            ... :341-411
   lambda == NULL selects the objective-only method (tail zero-filled, :295-297).
   vals[nnzh] in hess_structure! order:
   [obj k-block | obj FE | cons k-block | cons FE], lower triangle on global ids
   (src/fill_matrix_hessian_functions.jl:50-94); the obj FE block is absent for a NoFETerm
   objective (pdenlp_info.nnzh_fe_obj == 0).  A cons FE block that is structurally zero (no
   lambda, or a residual that is affine in the FE fields) is zero-filled without being computed;
   with PDENLP_MEM_HOST it is filled in the caller's buffer by host threads and not transferred. */
int pdenlp_hess_coord(pdenlp_model* m, const double* x, const double* lambda,
                      double obj_weight, double* vals);
/* objective part only, vals[nnzh_obj]: all there is for an AffineFEOperator model, whose
   constraint-Hessian structure is empty (src/hessian_struct_nnzh_functions.jl:70-72) */
int pdenlp_hess_coord_obj(pdenlp_model* m, const double* x, double obj_weight, double* vals);

/* jprod_nln!(nlp, x, v, Jv)            src/GridapPDENLPModel.jl:473-487
   jtprod_nln!(nlp, x, v, Jtv)          src/GridapPDENLPModel.jl:505-519
   matrix-free: same result as coo_prod! on the jac_coord! values. */
int pdenlp_jprod(pdenlp_model* m, const double* x, const double* v, double* Jv);
int pdenlp_jtprod(pdenlp_model* m, const double* x, const double* v, double* Jtv);

/* hprod!(nlp, x, v, Hv; obj_weight)         src/GridapPDENLPModel.jl:303-324
   hprod!(nlp, x, lambda, v, Hv; obj_weight) src/GridapPDENLPModel.jl:413-431
   lambda == NULL selects the objective-only method (zeros when obj_weight==0). */
int pdenlp_hprod(pdenlp_model* m, const double* x, const double* lambda,
                 double obj_weight, const double* v, double* Hv);

/* jac_nln_structure!(nlp, rows, cols)  src/GridapPDENLPModel.jl:541-550
   hess_structure!(nlp, rows, cols)     src/GridapPDENLPModel.jl:330-339
   1-based Int64, same memory space as the callback vectors.  Under Julia the
   arrays Gridap computed in the constructor stay authoritative; these are used
   to cross-check them and, in this container, as the structure source. */
int pdenlp_jac_structure(pdenlp_model* m, int64_t* rows, int64_t* cols);
int pdenlp_hess_structure(pdenlp_model* m, int64_t* rows, int64_t* cols);

/* ---- COO -> compressed sparse, duplicates summed -----------------------
   What consumes the COO streams next: NLPModels jac()/hess() = sparse(rows, cols, vals), and the
   linear API of an AffineFEOperator model, whose jac_lin_structure!/jac_lin_coord! return
   findnz(get_matrix(op)) (src/jacobian_functions.jl:11-21, src/GridapPDENLPModel.jl:597-608): the
   assembled matrix in CSC order, rows ascending inside a column.  The plan is built once from the
   static structure (stable device sort); apply() is a deterministic gather-sum (no atomics), each
   entry adding its contributions in COO (= cell) order. */
#define PDENLP_ORDER_CSC 0
#define PDENLP_ORDER_CSR 1
typedef struct pdenlp_compaction pdenlp_compaction; /* opaque */
/* rows/cols: 1-based Int64 COO structure (host pointers, or device pointers when
   rows_cols_on_device != 0); the plan lives on the current CUDA device. */
int pdenlp_compaction_create(int64_t nnz, const int64_t* rows, const int64_t* cols, int64_t nrows, int64_t ncols,
                             int32_t order, int32_t rows_cols_on_device, pdenlp_compaction** out);
int pdenlp_compaction_destroy(pdenlp_compaction* p);
int pdenlp_compaction_nnz(const pdenlp_compaction* p, int64_t* nnz_out);
/* compressed structure: rows/cols (1-based, nnz_out each) and/or the 1-based pointer array
   (colptr for CSC: ncols+1 entries, rowptr for CSR: nrows+1); any of them may be NULL */
int pdenlp_compaction_structure(pdenlp_compaction* p, int64_t* rows, int64_t* cols, int64_t* major_ptr, int32_t on_device);
/* out_vals[e] = sum of the COO values mapped to compressed entry e (device pointers) */
int pdenlp_compaction_apply(pdenlp_compaction* p, const double* coo_vals, double* out_vals, void* cuda_stream);

/* Free-dof ranges this handle's cells touch, 0-based half-open, in the concatenated numbering of each space:
   ranges = {y_lo, y_hi, u_lo, u_hi}.  A PDENLP_MEM_HOST model stages only x[0:p], x[p+y_lo : p+y_hi],
   x[p+nfree_y+u_lo : p+nfree_y+u_hi] (and lambda[y_lo : y_hi]) of the caller's vectors — a cell slab of a sharded model
   uploads its own rows of the mesh instead of the whole vector (x and lambda are replicated). */
int pdenlp_input_ranges(const pdenlp_model* m, int64_t ranges[4]);

/* ---- several GPUs behind one handle (cells are independent units; one contiguous, 32-cell aligned cell
   slab per device).  One process, one host thread; every device gets its own pdenlp_model over its slab, its own
   stream and its own device copies of the dof ranges of x / lambda / v its cells touch.  All vector arguments are
   HOST pointers (replicated x, global outputs — what a Julia solver holds); page-locked memory (pdenlp_host_alloc)
   lets the devices' copies overlap.
     jac_coord / hess_coord: no reduction — device r writes its slice of every FE segment straight to its place in the
       global COO stream (the slabs' slices concatenate in device order); the dense kappa blocks are slab partials, summed
       on the host in device order over the rows each slab touches.
     obj / grad / cons / jprod / jtprod / hprod: slab partials; only interface dofs (rows touched by two slabs) and the
       kappa entries are summed, on the host, in device order — deterministic, and no device-to-device collective is
       needed because the result goes to host memory anyway.
   Same status codes and error string as the single-device entry points. */
typedef struct pdenlp_sharded pdenlp_sharded; /* opaque */
/* devices == NULL: CUDA devices 0 .. ndev-1.  desc->device and desc->memspace are ignored. */
int pdenlp_sharded_create(const pdenlp_desc* desc, int32_t ndev, const int32_t* devices, pdenlp_sharded** out);
int pdenlp_sharded_destroy(pdenlp_sharded* s);
/* sizes of the GLOBAL problem (nnzj / nnzh are the lengths of the global COO streams) */
int pdenlp_sharded_get_info(const pdenlp_sharded* s, pdenlp_info* info);
/* cells [c0, c1) owned by device slot k */
int pdenlp_sharded_slab(const pdenlp_sharded* s, int32_t k, int64_t* c0, int64_t* c1);
int pdenlp_sharded_obj(pdenlp_sharded* s, const double* x, double* f);
int pdenlp_sharded_grad(pdenlp_sharded* s, const double* x, double* g);
int pdenlp_sharded_cons(pdenlp_sharded* s, const double* x, double* c);
int pdenlp_sharded_jac_coord(pdenlp_sharded* s, const double* x, double* vals);
int pdenlp_sharded_hess_coord(pdenlp_sharded* s, const double* x, const double* lambda, double obj_weight, double* vals);
int pdenlp_sharded_jprod(pdenlp_sharded* s, const double* x, const double* v, double* Jv);
int pdenlp_sharded_jtprod(pdenlp_sharded* s, const double* x, const double* v, double* Jtv);
int pdenlp_sharded_hprod(pdenlp_sharded* s, const double* x, const double* lambda, double obj_weight, const double* v, double* Hv);
int pdenlp_sharded_jac_structure(pdenlp_sharded* s, int64_t* rows, int64_t* cols);
int pdenlp_sharded_hess_structure(pdenlp_sharded* s, int64_t* rows, int64_t* cols);

/* bench support: number of kernel launches issued by this handle so far, and
   the CUDA-event time (ms) of the most recent callback on the handle's stream */
int pdenlp_launch_count(const pdenlp_model* m, int64_t* launches);
int pdenlp_last_kernel_ms(pdenlp_model* m, float* ms);
/* template path of jac_coord! / hess_coord!: stats[0] = chunks of consecutive template cells,
   stats[1] = template cells (copied from the image), stats[2] = generic cells (evaluated one by one), stats[3] = images
   built so far; all 0 when the model evaluates every cell (no template path, small slab, PDENLP_FLAG_NO_TEMPLATES) */
int pdenlp_template_stats(const pdenlp_model* m, int64_t stats[4]);
/* jac_coord! / hess_coord! of small models with kappa blocks are chains of short launches; from the second call with
   the same (x, lambda, vals, obj_weight, stream) on they are replayed as one CUDA graph (device memory space only; the
   contents of x / lambda are read at replay time; PDENLP_NO_GRAPHS=1 disables).  Number of replays so far: */
int pdenlp_graph_replays(const pdenlp_model* m, int64_t* replays);
int pdenlp_set_timing(pdenlp_model* m, int32_t enabled);

#ifdef __cplusplus
}
#endif
#endif /* PDENLP_CUDA_H */
