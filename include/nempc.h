/*
 * nempc.h -- C ABI of the B200-native NLP-evaluation hot path of pyNeuralEMPC.
 *
 * The reference has no FFI: its boundary for this path is four duck-typed Python classes plus the
 * cyipopt callback protocol.  Each entry point below states which reference
 * interface it replaces (paths relative to /root/reference/pyNeuralEMPC/).  The Python mirror of those
 * classes lives in pyneuralempc_b200/ and binds this library with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns 0 (NEMPC_OK) or a negative NEMPC_E* code; nothing throws across the ABI;
 *     nempc_last_error() gives the message (pass NULL for errors of nempc_create itself).
 *   - the caller owns all I/O buffers; the handle owns device weights, index tables and staging.
 *   - one CUDA stream per call; calls on ONE handle are not re-entrant (the reference's callbacks are
 *     single-threaded: optimizer/ipopt.py:189); independent handles may be used from different threads/GPUs.
 *   - there is NO CPU fallback: every compute entry point fails with NEMPC_ECUDA when no device is usable.
 *
 * Problem layout (identical to the reference)
 *   decision vector z = [x_1 .. x_H | u_0 .. u_{H-1}], n = H*(x_dim+u_dim)      optimizer/ipopt.py:20-28
 *   constraint row r = t*x_dim + p,  c_t = Phi(x_{t-1}, u_t) - x_t,  m = H*x_dim  integrator/discret.py:13-30
 *   batches are row-major: z (B,n), x0 (B,x_dim), lambda (B,m), resid (B,m), jac_vals (B,nnz_jac),
 *   hes_vals (B,nnz_hes), obj (B), grad (B,n); element type = desc.io_dtype.
 *   Jacobian values follow nempc_structure() (row-major order of the non-zeros of the dense matrix of
 *   integrator/discret.py:32-58 / rk4.py:113-178); Hessian values follow
 *   np.nonzero(np.tril(objective_map + integrator_map)) of optimizer/ipopt.py:55-62.
 */
#ifndef NEMPC_H_
#define NEMPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEMPC_MAX_LAYERS 8   /* dense layers including the linear output layer */
#define NEMPC_MAX_D 16       /* x_dim + u_dim */
#define NEMPC_MAX_EXO 64     /* tvp_dim + p_dim */

enum { NEMPC_OK = 0, NEMPC_EINVAL = -1, NEMPC_ECUDA = -2, NEMPC_ENOMEM = -3, NEMPC_ESTATE = -4, NEMPC_EUNSUPPORTED = -5 };
enum { NEMPC_F32 = 0, NEMPC_F64 = 1 };
enum { NEMPC_INTEG_DISCRETE = 0,   /* x_{t-1} + f - x_t        integrator/discret.py:13-30 */
       NEMPC_INTEG_UNITY = 1,      /* f - x_t                  integrator/unity.py:15-32   */
       NEMPC_INTEG_RK4 = 2 };      /* classical RK4, ZOH on u  integrator/rk4.py:57-83      */
enum { NEMPC_ACT_TANH = 0, NEMPC_ACT_SIGMOID = 1, NEMPC_ACT_SOFTPLUS = 2, NEMPC_ACT_RELU = 3 };
enum { NEMPC_KERNEL_AUTO = 0, NEMPC_KERNEL_GENERIC = 1, NEMPC_KERNEL_FAST = 2,
       NEMPC_KERNEL_TC = 3 };      /* tensor-core kernels: f32 tanh networks whose hidden layers are all 256 / 128 / 64 / 32 wide (tcgen05, split-f16 operands),
                                      f64 tanh networks whose hidden layers are all 128 / 64 wide (FP64 tensor cores, mma.sync.m8n8k4.f64); nempc_create
                                      returns NEMPC_EUNSUPPORTED when no instantiation matches */

typedef struct nempc_desc {
    int32_t x_dim, u_dim;                /* model/base.py:4-9 */
    int32_t horizon;                     /* H, integrator/base.py:19 */
    int32_t n_layers;                    /* Keras Dense layers, hidden (activation) + 1 linear output */
    int32_t widths[NEMPC_MAX_LAYERS];    /* fan-out of each layer; widths[n_layers-1] == x_dim */
    int32_t activation;                  /* NEMPC_ACT_* of the hidden layers */
    int32_t integrator;                  /* NEMPC_INTEG_* */
    double  dt;                          /* RK4 step (integrator/rk4.py:48); ignored otherwise */
    int32_t compute_dtype;               /* network + chain-rule arithmetic: NEMPC_F32 (reference: Keras f32) or F64 */
    int32_t io_dtype;                    /* element type of every I/O buffer (reference: f64, ipopt.py) */
    int32_t device;                      /* CUDA ordinal */
    int32_t kernel;                      /* NEMPC_KERNEL_* (AUTO: register-resident kernel for the small LV class, tensor-core kernel for 128- / 64-wide nets, else generic) */
    int32_t tvp_dim, p_dim;              /* time-varying / constant model inputs appended to (x, u): the network input is [x, u, tvp, p]
                                          * (model/tensorflow.py:39-47, model/base.py:4-9); 0 = none.  Layer 0 then has x+u+tvp+p input rows.
                                          * Served by the generic and the tensor-core kernels. */
} nempc_desc;

typedef struct nempc_handle nempc_handle;

const char* nempc_version(void);
/* ABI generation of the library (NEMPC_ABI_VERSION at its build) and the sizes of the two structs it was compiled with; a binding
 * compares them with its own after dlopen (a stale libnempc.so would otherwise misread nempc_desc / nempc_solver_opts). */
#define NEMPC_ABI_VERSION 4
int32_t nempc_abi_info(int32_t* desc_bytes, int32_t* solver_opts_bytes);
/* hash of the CUDA sources the binary was built from ("unknown" for a build outside pyneuralempc_b200/build.py) */
const char* nempc_source_hash(void);
const char* nempc_last_error(const nempc_handle* h);

/* ---- lifetime --------------------------------------------------------------------------------------- */
/* replaces the construction chain Model -> Integrator -> ObjectiveFunc -> IpoptProblem
 * (model/tensorflow.py:9-29, integrator/rk4.py:48-54, optimizer/ipopt.py:8-18). */
int nempc_create(const nempc_desc* desc, nempc_handle** out);
int nempc_destroy(nempc_handle* h);

/* Keras kernel layout W[in][out] row-major, y = x @ W + b (examples/lotka_volterra/nn_model.h5); HOST pointers,
 * always double (float32 weights are exactly representable). */
int nempc_set_weights(nempc_handle* h, int32_t layer, const double* W, const double* b);

/* separable cost f(z) = sum_i lin_i z_i + quad_i (z_i - ref_i)^2 -- the families used with
 * objective/jax.py:28-57 (run.py:83-84 linear in u, test.py:59-60 squared set-point, diagonal tracking).
 * HOST pointers of length n, NULL = zeros.  Changes the Hessian pattern where quad != 0 on x_H. */
int nempc_set_objective(nempc_handle* h, const double* lin, const double* quad, const double* ref);

/* Exogenous model inputs of the following nempc_eval* / nempc_solve calls: what NMPC.next hands down as `tvp` (H, tvp_dim) and `p`
 * (p_dim,) (controller.py:65-113 -> ProblemFactory.set_tvp / set_p -> Model.forward(x, u, p=, tvp=), model/tensorflow.py:39-47).
 * They are inputs of the network but not decision variables: the Jacobian / Hessian keep their (x, u) shape, exactly the column
 * slicing of model/tensorflow.py:65-66, 97-98.  The reference's `p` branch is mis-shaped (tensorflow.py:44-45, its own TODO); the
 * intended meaning -- the same p row appended to every sample of a problem -- is implemented.
 *   tvp: tvp_rows x tvp_dim doubles; tvp_rows == H shares one table between all problems of a batch, otherwise row b*H + t belongs
 *        to step t of problem b (nempc_model_eval: row i belongs to sample i).
 *   p:   p_rows x p_dim doubles; p_rows == 1 shares one row between all problems, otherwise row b belongs to problem b.
 * HOST or DEVICE pointers (copied into buffers owned by the handle; synchronises the device).  NULL for a kind the model lacks. */
int nempc_set_exogenous(nempc_handle* h, int64_t tvp_rows, const double* tvp, int64_t p_rows, const double* p);

/* ---- sparsity (host only, no device needed) ----------------------------------------------------------- */
/* replaces Integrator.hessianstructure / JAXObjectifFunc.hessianstructure / IpoptProblem.hessianstructure
 * (integrator/base.py:83-115, objective/jax.py:59-90, optimizer/ipopt.py:55-62) with the analytic pattern,
 * and adds the Jacobian pattern the reference leaves dense (optimizer/ipopt.py:88-96).
 * quad_mask: n bytes, non-zero where the objective has a diagonal Hessian entry (NULL = none). */
int nempc_structure_counts(int32_t horizon, int32_t x_dim, int32_t u_dim, const uint8_t* quad_mask,
                           int64_t* nnz_jac, int64_t* nnz_hes);
int nempc_structure_fill(int32_t horizon, int32_t x_dim, int32_t u_dim, const uint8_t* quad_mask,
                         int32_t* jac_row, int32_t* jac_col, int32_t* hes_row, int32_t* hes_col);
int nempc_dims(const nempc_handle* h, int64_t* n, int64_t* m, int64_t* nnz_jac, int64_t* nnz_hes);
int nempc_structure(const nempc_handle* h, int32_t* jac_row, int32_t* jac_col, int32_t* hes_row, int32_t* hes_col);

/* ---- the hot path ------------------------------------------------------------------------------------- */
/* One NLP evaluation of B independent problems at one iterate; replaces, per problem,
 *   IpoptProblem.constraints (ipopt.py:44-52)  -> resid      Integrator.forward  (discret.py:13-30, rk4.py:57-83)
 *   IpoptProblem.jacobian    (ipopt.py:88-96)  -> jac_vals   Integrator.jacobian (discret.py:32-58, rk4.py:113-178)
 *   IpoptProblem.hessian     (ipopt.py:66-86)  -> hes_vals   Integrator.hessian  (discret.py:61-81, rk4.py:181-285)
 *                                                            + obj_factor * ObjectiveFunc.hessian (objective/jax.py:43-57)
 *   IpoptProblem.objective / gradient (ipopt.py:30-42) -> obj, grad
 * DEVICE pointers, element type io_dtype.  Any output may be NULL (skipped); hes_vals needs lambda.
 * obj_factor: device array of B values, or NULL to use obj_factor_scalar for every problem.
 * stream: a cudaStream_t (NULL = default stream).  Asynchronous. */
int nempc_eval(nempc_handle* h, int64_t B, const void* z, const void* x0, const void* lambda,
               const void* obj_factor, double obj_factor_scalar,
               void* resid, void* jac_vals, void* hes_vals, void* obj, void* grad, void* stream);

/* Same call with HOST buffers (the cyipopt callback situation); synchronises before it returns.
 *   - page-locked (mapped) buffers and a small call (<= 256 KB moved): ZERO-COPY, the kernels read / write the host buffers directly;
 *   - otherwise the batch is cut in chunks that pipeline H2D | kernels | D2H on three streams through device staging owned by the
 *     handle; with page-locked buffers the pipeline of an unchanged argument set is captured once and replayed as a CUDA graph
 *     (nempc_set_weights / nempc_set_objective invalidate it);
 *   - pageable buffers work too (the copies are then synchronous in the driver). */
int nempc_eval_host(nempc_handle* h, int64_t B, const void* z, const void* x0, const void* lambda,
                    const void* obj_factor, double obj_factor_scalar,
                    void* resid, void* jac_vals, void* hes_vals, void* obj, void* grad);

/* Per-step blocks before sparse assembly: pred (B,H,x) = Phi(x_{t-1},u_t) without the -x_t term,
 * AB (B,H,x,d) = d Phi / d(x_{t-1},u_t), Hblk (B,H,x,d,d) per-output second derivatives (not lambda-contracted):
 * what Integrator.jacobian / Integrator.hessian scatter densely (rk4.py:159-176, 266-283).  DEVICE pointers. */
int nempc_eval_blocks(nempc_handle* h, int64_t B, const void* z, const void* x0,
                      void* pred, void* AB, void* Hblk, void* stream);

/* Raw network value / Jacobian / per-output Hessian of N stacked inputs zin (N,d):
 * replaces KerasTFModel.forward / .jacobian / .hessian per sample (model/tensorflow.py:49-109),
 * f (N,x), jac (N,x,d), hes (N,x,d,d).  DEVICE pointers; jac / hes may be NULL. */
int nempc_model_eval(nempc_handle* h, int64_t N, const void* zin, void* f, void* jac, void* hes, void* stream);

/* Stand-alone objective value / gradient of the separable cost (no handle needed):
 * replaces JAXObjectifFunc.forward / .gradient (objective/jax.py:28-41).  z (B,n) and obj (B) / grad (B,n) are
 * DEVICE arrays of io_dtype; lin / quad / ref are DEVICE arrays of n doubles.  obj or grad may be NULL. */
int nempc_objective_eval(int32_t io_dtype, int64_t B, int64_t n, const void* z, const double* lin, const double* quad,
                         const double* ref, void* obj, void* grad, void* stream);

/* ---- non-separable quadratic costs ------------------------------------------------------------------------------
 * JAXObjectifFunc takes any scalar function (objective/jax.py:28-57).  Beyond the separable family of nempc_set_objective the device
 * path evaluates GENERAL QUADRATIC costs f(z) = 1/2 z' P z + q' z + c with a sparse symmetric P: control-rate penalties
 * (u_{t+1} - u_t)' S (u_{t+1} - u_t), full-matrix stage weights (x_t - r_t)' Q (x_t - r_t), a full terminal weight on x_H, cross terms.
 *   nempc_quadform_eval   obj (B) / grad (B, n) of that cost; P as CSR over all n rows with BOTH triangles (p_ptr n+1, p_idx, p_val),
 *                         q (n) and the table pointers are DEVICE arrays; z / obj / grad DEVICE arrays of io_dtype.
 *   nempc_hessian_merge   sigma * P joins the Lagrangian Hessian (optimizer/ipopt.py:66-86) on the UNION pattern
 *                         np.nonzero(np.tril(objective_map + integrator_map)) (ipopt.py:55-62):
 *                         out_vals[b, s] = kern_vals[b, src_slot[s]] (src_slot[s] >= 0: the slot of the constraint part as nempc_eval
 *                         wrote it on a handle WITHOUT a device objective) + obj_factor_b * p_val[s].
 * Stateless, asynchronous on `stream`. */
int nempc_quadform_eval(int32_t io_dtype, int64_t B, int64_t n, const void* z, const int32_t* p_ptr, const int32_t* p_idx,
                        const double* p_val, const double* q, double c, void* obj, void* grad, void* stream);
int nempc_hessian_merge(int32_t io_dtype, int64_t B, int64_t nnz_kern, int64_t nnz_out, const void* kern_vals,
                        const int32_t* src_slot, const double* p_val, const void* obj_factor, double obj_factor_scalar,
                        void* out_vals, void* stream);

/* ---- rolling-window (NARX) models ---------------------------------------------------------------
 * Replaces the window logic of KerasTFModelRollingInput / DiffDiscretJaxModelRollingWindow (model/tensorflow.py:112-340,
 * model/jax.py:93-259: rolling_input, _gather_input, the projection matrices of jacobian / hessian) together with the dense slicing
 * the integrators apply to the model's arrays (integrator/discret.py:32-81, unity.py:34-81).  The network of step t reads the last
 * w = rolling_window states and controls, dw = w (x + u) inputs; its per-row value / Jacobian / per-output Hessian come from
 * nempc_model_eval on a handle created with x_dim = x, u_dim = dw - x.  All data pointers are DEVICE pointers of io_dtype, the index
 * tables are DEVICE int32 / double arrays built once by the caller (pyneuralempc_b200/rolling.py builds them from the closed-form
 * band structure).  No handle: both calls are stateless and asynchronous on `stream`.
 *   gather:    zin[b, r] = z[b, gidx[r]] if gidx[r] >= 0 else aux[b, -1 - gidx[r]],  r < rows = H * dw,
 *              aux (B, naux) = [x0 | prev_x | prev_u] (KerasTFModelRollingInput.set_prev_data, tensorflow.py:174-185)
 *   assemble:  resid[b, r]   = f[b, r] - z[b, r] (+ the gathered x_{t-1} entry resid_base[r]; INT32_MIN = none, unity integrator)
 *              jac_vals[b,s] = jac_add[s] (+ J[b].flat[jac_src[s]] when jac_src[s] >= 0)
 *              hes_vals[b,s] = obj_factor_b * hes_obj[s] + sum_{e in [hes_ptr[s], hes_ptr[s+1])} sum_p lambda[b, t x + p] Hs[b, t, p].flat[ab],
 *                              hes_src[e] = t dw^2 + ab                    (fixed summation order: deterministic)
 * Any of resid / jac_vals / hes_vals may be NULL (skipped). */
int nempc_rolling_gather(int32_t io_dtype, int64_t B, int32_t n, int32_t naux, int32_t rows, const int32_t* gidx,
                         const void* z, const void* aux, void* zin, void* stream);
int nempc_rolling_assemble(int32_t io_dtype, int64_t B, int32_t H, int32_t x, int32_t dw, int32_t n, int32_t naux,
                           int64_t nnz_jac, int64_t nnz_hes, const int32_t* resid_base, const int32_t* jac_src,
                           const double* jac_add, const int32_t* hes_ptr, const int32_t* hes_src, const double* hes_obj,
                           const void* z, const void* aux, const void* f, const void* J, const void* Hs, const void* lambda,
                           const void* obj_factor, double obj_factor_scalar, void* resid, void* jac_vals, void* hes_vals,
                           void* stream);

/* ---- batched on-device NMPC solver ----------------------------------------------------------
 * Replaces, for a batch of B independent problems, the solver call of Ipopt.solve / Slsqp.solve
 * (optimizer/ipopt.py:162-189, optimizer/slsqp.py:172-173): a primal-dual interior-point method whose Newton step is a
 * Riccati sweep over the block-tridiagonal KKT system, consuming the Jacobian / Hessian value arrays of nempc_eval in
 * place.  Needs io_dtype = F64 and nempc_set_objective.  lb / ub: HOST arrays of n doubles in DomainConstraint order
 * (constraints.py:26-30), +-inf = unbounded.  x0 (B,x), z (B,n), lambda (B,m), kkt_error (B): DEVICE doubles;
 * status / iterations (B): DEVICE int32.  z is the initial guess when use_init != 0 (else [x0 tiled | zeros],
 * optimizer/ipopt.py:149) and receives the solution.  status: 0 converged, 1 iteration limit, 2 failed (non-finite step:
 * the problem is infeasible or unbounded; the reference returns Optimizer.FAIL there).  Synchronises the stream. */
typedef struct nempc_solver_opts {
    int32_t max_iter, max_backtrack;
    double tol, mu_init, mu_min, kappa_eps, kappa_mu, theta_mu, tau_min, bound_push, eta, reg_init, reg_max;
} nempc_solver_opts;
int nempc_solver_defaults(nempc_solver_opts* opts);
int nempc_solve(nempc_handle* h, int64_t B, const void* x0, const double* lb, const double* ub, void* z, int32_t use_init,
                void* lambda, int32_t* status, int32_t* iterations, double* kkt_error, const nempc_solver_opts* opts,
                int32_t* outer_iterations, void* stream);
/* Statistics of the last nempc_solve on this handle.  used_graph: 1 when the whole iteration loop ran on the device as ONE CUDA graph
 * (two nested WHILE nodes driven by cudaGraphSetConditional: no host synchronisation between the kernels of an iteration; environment
 * NEMPC_SOLVE_GRAPH = 0 never / 1 from the second solve with unchanged batch size, options and workspace (default) / 2 from the first),
 * 0 when the host issued the kernels.  Both produce the same bits.  unaccepted_steps: steps over all problems that were taken although no
 * line-search trial within max_backtrack halvings passed the Armijo test (the last halved step is applied; a large count with status 1
 * says the tolerance is below what the network's float32 arithmetic resolves). */
int nempc_solve_stats(const nempc_handle* h, int32_t* used_graph, int64_t* unaccepted_steps);

/* ---- introspection -------------------------------------------------------------------------------------- */
int64_t nempc_launch_count(const nempc_handle* h);   /* kernels launched by this handle so far */
const char* nempc_kernel_name(const nempc_handle* h); /* "fast_mlp2<...>" or "generic<...>" */
/* algorithmic FLOP per horizon step (SURVEY 8d formula) for the roofline report */
double nempc_flops_per_step(const nempc_handle* h);
/* sustained FMA throughput of the device's FP32 / FP64 pipe (TFLOP/s), measured with a register-resident
 * FMA loop for about `millis` ms: the denominator of the compute roofline. */
int nempc_measure_fma_peak(int32_t device, int32_t dtype, int32_t millis, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* NEMPC_H_ */
