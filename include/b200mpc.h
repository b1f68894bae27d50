/*
 * b200mpc.h -- C ABI of the B200-native batched MPC solve engine.
 *
 * The reference (nicolapiccinelli/libmpc, libmpc++ v0.7.1) has no FFI: the seam this library sits behind is the
 * C++ virtual interface mpc::IOptimizer<sizer> (include/mpc/IOptimizer.hpp:24-58) together with the setters of
 * mpc::LMPC<> that feed mpc::ProblemBuilder / mpc::LOptimizer.  Each entry point below names the reference
 * member it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - extern "C", opaque handle, plain pointers and sizes, int return code (0 = ok, <0 = B200MPC_E*), no
 *     exceptions cross the boundary.
 *   - All matrices are FP64.  A/B/C/Bd/Dd are ROW-major [rows x cols].  Every horizon matrix is STAGE-major:
 *     [ph][dim] (== the memory image of the reference's column-major Eigen mat<dim,ph>).
 *   - `per_instance` = 0: one copy shared by the whole batch; 1: `batch` consecutive copies.
 *   - `dev` = 0: pointer is host memory (copied with cudaMemcpyAsync on the handle's stream);
 *     `dev` = 1: pointer is device memory on the handle's device (copied device-to-device, no host involvement).
 *   - The engine solves `batch` independent LMPC problems per call.  mpc::LMPC<>::optimize() is batch == 1.
 */
#ifndef B200MPC_H
#define B200MPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MPC_OK 0
#define B200MPC_EINVAL (-1)  /* bad argument / dimension                                                     */
#define B200MPC_ECUDA (-2)   /* CUDA runtime error (b200mpc_last_error() holds the string)                   */
#define B200MPC_ENOGPU (-3)  /* no CUDA device: the engine has no CPU fallback                               */
#define B200MPC_ESTATE (-4)  /* call order violation (e.g. solve before set_model)                           */

typedef struct b200mpc_lmpc* b200mpc_lmpc_t;

/* mpc::MPCSize for LMPC<Tnx,Tnu,Tndu,Tny,Tph,Tch>   (include/mpc/LMPC.hpp:23-26, include/mpc/Dim.hpp:107-132) */
typedef struct b200mpc_lmpc_dims {
    int nx, nu, ndu, ny, ph, ch;
} b200mpc_lmpc_dims;

/* mpc::LParameters (include/mpc/Types.hpp:142-160) + the OSQP v0.6.3 defaults LOptimizer::run inherits through
 * osqp_set_default_settings (include/mpc/LMPC/LOptimizer.hpp:244-257).  b200mpc_lmpc_default_params fills the
 * reference defaults. */
typedef struct b200mpc_lmpc_params {
    int maximum_iteration;        /* Parameters::maximum_iteration = 100                                     */
    int enable_warm_start;        /* Parameters::enable_warm_start = false                                   */
    double alpha;                 /* 1.6                                                                     */
    double rho;                   /* 1e-6                                                                    */
    double eps_rel, eps_abs;      /* 1e-4                                                                    */
    double eps_prim_inf, eps_dual_inf; /* 1e-3                                                               */
    int adaptive_rho;             /* true                                                                    */
    int polish;                   /* true                                                                    */
    /* OSQP defaults libmpc does not override */
    double sigma;                 /* 1e-6                                                                    */
    double delta;                 /* 1e-6                                                                    */
    double adaptive_rho_tolerance;/* 5                                                                       */
    int scaling;                  /* 10                                                                      */
    int check_termination;        /* 25                                                                      */
    int adaptive_rho_interval;    /* 25: v0.6.3 derives it from wall-clock time; pinned      */
    int polish_refine_iter;       /* 3                                                                       */
    double time_limit;            /* Parameters::time_limit = 0 (off), seconds (LOptimizer.hpp:256): OSQP checks its run time
                                     (set-up included) at the top of every ADMM iteration; when it is exceeded the solve stops
                                     with OSQP_TIME_LIMIT_REACHED (-6), which LOptimizer::convertToResultStatus maps to UNKNOWN
                                     (LOptimizer.hpp:386-415).  Per controller, measured with the GPU's global timer from the
                                     moment its solve starts; honoured by the CTA-per-controller engine               */
} b200mpc_lmpc_params;

/* mpc::ResultStatus (include/mpc/Types.hpp:84-91) */
enum { B200MPC_SUCCESS = 0, B200MPC_MAX_ITERATION = 1, B200MPC_INFEASIBLE = 2, B200MPC_ERROR = 3, B200MPC_UNKNOWN = 4 };

const char* b200mpc_last_error(void);
int b200mpc_device_count(void);

void b200mpc_lmpc_default_params(b200mpc_lmpc_params* p);

/* LMPC::LMPC()/onSetup -> ProblemBuilder::onInit + LOptimizer::onInit (LMPC.hpp:51-61,728-735;
 * ProblemBuilder.hpp:88-172; LOptimizer.hpp:59-82): allocates device state for `batch` controllers with the
 * reference defaults (zero model/weights, infinite bounds, zero references). */
int b200mpc_lmpc_create(const b200mpc_lmpc_dims* dims, int batch, int device, b200mpc_lmpc_t* out);
int b200mpc_lmpc_destroy(b200mpc_lmpc_t h);
/* cudaStream_t to run on (0 = legacy default stream); passed as void* to keep cuda headers out of this file. */
int b200mpc_lmpc_set_stream(b200mpc_lmpc_t h, void* stream);

/* LMPC::setOptimizerParameters (LMPC.hpp:79-82) -> LOptimizer::setParameters (LOptimizer.hpp:100-108) */
int b200mpc_lmpc_set_params(b200mpc_lmpc_t h, const b200mpc_lmpc_params* p);

/* LMPC::setStateSpaceModel(A,B,C) (LMPC.hpp:493-500) -> ProblemBuilder::setStateModel (ProblemBuilder.hpp:184-211)
 * A[nx*nx], B[nx*nu], C[ny*nx] row-major. */
int b200mpc_lmpc_set_model(b200mpc_lmpc_t h, const double* A, const double* B, const double* C,
                           int per_instance, int dev);
/* LMPC::setDisturbances(Bd,Dd) (LMPC.hpp:518-525) -> ProblemBuilder::setExogenousInput (:222-236) */
int b200mpc_lmpc_set_disturbances(b200mpc_lmpc_t h, const double* Bd, const double* Dd, int per_instance, int dev);
/* LMPC::setObjectiveWeights(OWeightMat,UWeightMat,DeltaUWeightMat) (LMPC.hpp:306-313) -> setObjective (:247-263)
 * OW[ph*ny], UW[ph*nu], DUW[ph*nu] stage-major. */
int b200mpc_lmpc_set_weights(b200mpc_lmpc_t h, const double* OW, const double* UW, const double* DUW,
                             int per_instance, int dev);
/* LMPC::setStateBounds / setInputBounds / setOutputBounds matrix forms (LMPC.hpp:111-141) -> ProblemBuilder
 * (:378-432).  XMin/XMax[ph*nx], UMin/UMax[ch*nu] (the tail ph-ch is replicated as in :406-410), YMin/YMax[ph*ny].
 * Infinite bounds are IEEE +-inf exactly like mpc::inf (Types.hpp:228). */
int b200mpc_lmpc_set_state_bounds(b200mpc_lmpc_t h, const double* XMin, const double* XMax, int per_instance, int dev);
int b200mpc_lmpc_set_input_bounds(b200mpc_lmpc_t h, const double* UMin, const double* UMax, int per_instance, int dev);
int b200mpc_lmpc_set_output_bounds(b200mpc_lmpc_t h, const double* YMin, const double* YMax, int per_instance, int dev);
/* The reference's INTERNAL input-bound matrices minU/maxU [nu x ph] as they stand after any mix of the matrix setter (which
 * replicates column ch-1 into the tail, ProblemBuilder.hpp:397-413) and the per-index setter (which writes one column and does
 * NOT touch the tail, ProblemBuilder.hpp:469-477): UMin/UMax[ph*nu], all ph columns taken verbatim. */
int b200mpc_lmpc_set_input_bounds_full(b200mpc_lmpc_t h, const double* UMin, const double* UMax, int per_instance, int dev);
/* LMPC::setScalarConstraint(min,max,X,U,slice all) (LMPC.hpp:355-407) -> ProblemBuilder::setScalarConstraint
 * (:347-365).  SMin/SMax[ph], X[nx], U[nu].  As in the reference the multiplier applies to every stage. */
int b200mpc_lmpc_set_scalar_constraint(b200mpc_lmpc_t h, const double* SMin, const double* SMax, const double* X,
                                       const double* U, int per_instance, int dev);
/* LMPC::setReferences(outRefMat,cmdRefMat,deltaCmdRefMat) (LMPC.hpp:596-602) -> LOptimizer::setReferences
 * (LOptimizer.hpp:119-129).  yRef[ph*ny], uRef[ph*nu], duRef[ph*nu]. */
int b200mpc_lmpc_set_references(b200mpc_lmpc_t h, const double* yRef, const double* uRef, const double* duRef,
                                int per_instance, int dev);
/* LMPC::setExogenousInputs(uMeasMat) (LMPC.hpp:534-538) -> LOptimizer::setExogenousInputs (:161-165). uMeas[ph*ndu] */
int b200mpc_lmpc_set_exogenous_inputs(b200mpc_lmpc_t h, const double* uMeas, int per_instance, int dev);

/* LMPC::setSolverWarmStart / getSolverWarmStartPrimal/Dual (LMPC.hpp:677-722): primal[batch*n], dual[batch*m] in the
 * reference's variable/row order.  n = (ph+1)(nx+nu)+ph*nu, m = 2(ph+1)(nx+nu)+(ph+1)ny+ph*nu+(ph+1). */
int b200mpc_lmpc_set_warm_start(b200mpc_lmpc_t h, const double* primal, const double* dual, int dev);
int b200mpc_lmpc_get_warm_start(b200mpc_lmpc_t h, double* primal, double* dual, int dev);

/* IOptimizer::run(x0,u0) (IOptimizer.hpp:50, LOptimizer.hpp:189-368) for `batch` instances:
 * x0[batch*nx], u0[batch*nu].  Enqueues the copies + ONE kernel on the handle's stream; does not synchronise.
 * Results stay on the device until fetched. */
int b200mpc_lmpc_solve(b200mpc_lmpc_t h, const double* x0, const double* u0, int dev);

/* Closed loop on the device: the control loop every example wraps around optimize()
 * (examples/quadrotor_ex.cpp, ugv_ex.cpp:143-166: optimize -> apply cmd -> step the plant -> repeat) for `steps` control
 * steps without leaving the GPU: step k solves from (x_k, u_{k-1}), applies u_k = cmd and advances the plant
 * x_{k+1} = Ap x_k + Bp u_k.  Ap[nx*nx] / Bp[nx*nu] row-major (plant_per_instance: batch copies); NULL = the controller's
 * own model.  With enable_warm_start every solve starts from the previous optimum exactly as LOptimizer::run does
 * (LOptimizer.hpp:268-281).  Outputs: traj_x[(steps+1)*batch*nx] (x_0..x_steps), traj_u[steps*batch*nu],
 * traj_status / traj_iters[steps*batch] (may be NULL).  2 kernel launches per step, no host round trip; synchronises at the end. */
int b200mpc_lmpc_closed_loop(b200mpc_lmpc_t h, const double* x0, const double* u0, int steps, const double* Ap,
                             const double* Bp, int plant_per_instance, double* traj_x, double* traj_u,
                             int32_t* traj_status, int32_t* traj_iters, int dev);

/* One plant step of that loop on its own: x <- A x + B cmd with the controller's model (x_dev [batch*nx], updated in place) and
 * u_out_dev [batch*nu] <- cmd, asynchronous on the handle's stream (solve / advance pairs enqueue back to back). */
int b200mpc_lmpc_advance(b200mpc_lmpc_t h, double* x_dev, double* u_out_dev);

/* mpc::Result<nu> fields (Types.hpp:168-182), one entry per instance.  Any pointer may be NULL.
 *   cmd[batch*nu], cost[batch], status[batch] (ResultStatus), solver_status[batch] (OSQP status_val),
 *   is_feasible[batch] (0/1), iterations[batch], rho_updates[batch], status_polish[batch]
 * `dev`=0 copies to host and synchronises the stream; `dev`=1 copies device-to-device asynchronously. */
int b200mpc_lmpc_get_result(b200mpc_lmpc_t h, double* cmd, double* cost, int32_t* status, int32_t* solver_status,
                            int32_t* is_feasible, int32_t* iterations, int32_t* rho_updates,
                            int32_t* status_polish, int dev);
/* mpc::OptSequence (Types.hpp:184-199) as filled by LOptimizer::run (:305-338): state[batch*(ph+1)*nx],
 * input[batch*(ph+1)*nu], output[batch*(ph+1)*ny], row-major [(ph+1) x dim] per instance. */
int b200mpc_lmpc_get_sequence(b200mpc_lmpc_t h, double* state, double* input, double* output, int dev);
/* Device pointer of the command block cmd[batch*nu] (for a fused NCCL all-gather on the same stream). */
int b200mpc_lmpc_cmd_device_ptr(b200mpc_lmpc_t h, double** cmd_dev);

/* ---- multi-GPU: the batch shards across ranks (one process per GPU, rank r owns a contiguous block of
 * controllers) with no data-path collective; the ONE exchange step is an all-gather of the command blocks.  The reference is a
 * single-threaded, single-controller library and has no counterpart; the contract is SURVEY.md 8b's b200mpc_comm_init /
 * _allgather_cmd.  NCCL is loaded at run time (the copy already in the process when there is one).
 *   b200mpc_comm_unique_id: rank 0 creates the 128-byte ncclUniqueId, the caller broadcasts it by any means;
 *   b200mpc_comm_init_rank: every rank joins (ncclCommInitRank on `device`);   b200mpc_comm_init: wrap an existing ncclComm_t;
 *   b200mpc_lmpc_allgather_cmd: cmd[batch*nu] of every rank -> cmd_all_dev[nranks*batch*nu] (device memory, rank-major),
 *   enqueued on the handle's stream directly behind the solve (send buffer = the kernel's output block; no host sync). */
typedef struct b200mpc_comm* b200mpc_comm_t;
int b200mpc_comm_unique_id(void* id128);
int b200mpc_comm_init_rank(int nranks, int rank, const void* id128, int device, b200mpc_comm_t* out);
int b200mpc_comm_init(void* nccl_comm, int nranks, int rank, b200mpc_comm_t* out);
int b200mpc_comm_destroy(b200mpc_comm_t c);
int b200mpc_comm_size(b200mpc_comm_t c, int* nranks, int* rank);
int b200mpc_lmpc_allgather_cmd(b200mpc_lmpc_t h, b200mpc_comm_t c, double* cmd_all_dev);

/* Engine introspection used by bench.py / profiles: resident warp slots, workspace bytes per slot, kernel
 * launches issued so far, algorithmic FP64 flop estimate of the last solve (sum over instances). */
int b200mpc_lmpc_info(b200mpc_lmpc_t h, int* warp_slots, size_t* workspace_bytes_per_slot, long long* launches);
/* Solve engine.  0 = automatic; 2 = one thread block per controller with its whole state (factor, vectors, iterates) in shared
 * memory, a persistent grid of one block per SM (the default whenever the vectors fit: batch 1 -- a single mpc::LMPC<> object --
 * up to any batch); 1 = one warp per controller, state streamed from HBM through TMA rings (any size).  cta_threads: 0 (auto), 256
 * or 512.  Results do not depend on the engine (tests/test_gpu_lmpc_properties.py).  Environment: B200MPC_ENGINE, B200MPC_CTA_THREADS. */
int b200mpc_lmpc_set_engine(b200mpc_lmpc_t h, int engine, int cta_threads);
int b200mpc_lmpc_get_engine(b200mpc_lmpc_t h, int* engine, int* threads_per_cta, int* factor_in_shared_memory);
/* Override launch geometry of engine 1 (0 = auto): warps per CTA and CTAs per SM of the persistent warp-per-controller kernel
 * (a non-zero value selects engine 1). */
int b200mpc_lmpc_set_launch(b200mpc_lmpc_t h, int warps_per_cta, int ctas_per_sm);
/* How the persistent warps draw instances.  GANG (default): the warps of a CTA draw and start their instances together
 * and wait for the slowest before drawing again -- they run the same phase of the program at the same time and share its
 * instruction-cache footprint (measured: 95k vs 48k solves/s at batch 32768, profiles/r01_icache.md).  FREE: every warp
 * draws on its own as soon as it is done (no waiting, but the phases interleave and the kernel becomes instruction-fetch
 * bound).  Results are identical. */
enum { B200MPC_SCHEDULE_FREE = 0, B200MPC_SCHEDULE_GANG = 1 };
int b200mpc_lmpc_set_schedule(b200mpc_lmpc_t h, int schedule);
/* History ordering (default on): from the second solve of a handle on, instances are drawn in the order of the iteration
 * counts of their previous solve, longest first.  Consecutive MPC steps of a controller take similar iteration counts, so
 * the members of a gang finish together and the long solves start first.  Affects scheduling only, never results.
 * (Experiment knobs: the environment variables B200MPC_GANG=<0|1|k> and B200MPC_HISTORY_ORDER=<0|1>, read when a handle is
 * created, set the initial value of these two switches.) */
int b200mpc_lmpc_set_history_order(b200mpc_lmpc_t h, int enable);
/* Profiling aid: first call (out_host ignored) enables per-instance phase cycle counters; later calls copy
 * batch*8 counters [setup, factorize, admm sweeps, info, polish prep, polish factor, polish solve, unpack] to the host. */
int b200mpc_lmpc_profile(b200mpc_lmpc_t h, long long* out_host);
int b200mpc_sync(b200mpc_lmpc_t h);

/* ---- set-up helper: mpc::discretization<nx,nu>(A,B,Ts,Ad,Bd) (include/mpc/Utils.hpp:23-47) for `batch`
 * systems in one launch -- zero-order-hold c2d through exp([[A,B],[0,0]] Ts).  A[nx*nx], B[nx*nu] row-major (one copy, or
 * `batch` copies with model_per_instance), Ts one value or `batch` values; outputs Ad[batch*nx*nx], Bd[batch*nx*nu]. */
int b200mpc_c2d(int nx, int nu, int batch, const double* A, const double* B, int model_per_instance, const double* Ts,
                int ts_per_instance, double* Ad, double* Bd, int dev, void* stream);

/* ---- NLMPC: batched evaluation of everything libmpc++ hands to its NLP solver for a decision vector z --------------
 * (SURVEY.md kernel K5).  Replaces, for `batch` controllers at once, the four NLopt callbacks of NLOptimizer
 * (include/mpc/NLMPC/NLOptimizer.hpp:760-997): Objective::evaluate (NLMPC/Objective.hpp:91-265),
 * Constraints::evaluateStateModelEq (NLMPC/Constraints.hpp:325-356,490-628,844-905) and Constraints::evaluateIneq
 * (NLMPC/Constraints.hpp:211-316,641-721), after Mapping::unwrapVector (NLMPC/Mapping.hpp:174-257).
 * The reference's std::function callbacks (IDimensionable.hpp:94-149) cannot run on the device, so the model / cost /
 * constraints are device functors selected by `system`; the reference's three example systems are built in.
 *   z[batch*nz] with nz = ph*nx + ch*nu + 1 ([X_1..X_ph ; U_1..U_ch ; slack], Mapping.hpp:196-201), x0[batch*nx],
 *   params[(batch or 1)*nparam]; outputs (any may be NULL): fval[batch], grad[batch*nz], ceq[batch*ph*nx],
 *   Jeq[batch*ph*nx*nz], cin[batch*nineq], Jin[batch*nineq*nz]; Jacobians row-major like the reference's. */
enum { B200MPC_SYS_VANDERPOL = 0,  /* examples/vanderpol_ex.cpp: params [Ts]                                   */
       B200MPC_SYS_OSCNET4 = 1,    /* examples/networked_oscillators_ex.cpp with N=4: params [Ts, mu, k]         */
       B200MPC_SYS_OSCNET6 = 2,    /* the shipped N=6                                                           */
       B200MPC_SYS_UGV = 3 };      /* examples/ugv_ex.cpp: params [Ad(16) Bd(8) v_pref(2) obs0(x,y,r) obs1(x,y,r)] */
int b200mpc_nlmpc_system_dims(int system, int ph, int* nx, int* nu, int* nparam, int* nineq);
/* number of user equality constraints Teq of the system (NLMPC<...,Tineq,Teq>, NLMPC.hpp:26-30); 0 for the built-ins */
int b200mpc_nlmpc_system_neq(int system, int ph, int* neq);
/* Tny of the system and whether it defines an output map (NLMPC::setOutputFunction, NLMPC.hpp:202-215) */
int b200mpc_nlmpc_system_ny(int system, int ph, int* ny, int* has_output_map);

/* ---- user-defined systems: NLMPC::setStateSpaceFunction / setObjectiveFunction / setIneqConFunction / setEqConFunction
 * (NLMPC.hpp:165,139,228,261; callback typedefs IDimensionable.hpp:94-149).  The reference takes host std::function
 * callbacks; the batched engine takes the same four pieces as CUDA source, compiled at run time (NVRTC, sm_100a) straight
 * into its kernels.  `cuda_source` must define, at global scope, a struct named `type_name` with
 *     static constexpr int nx, nu, ny, nparam;                 // dimensions, number of doubles in `params`
 *     static constexpr bool continuous;                        // true: dx/dt = f (trapezoidal collocation, needs Ts)
 *     __device__ static double Ts(const double* p);            // sampling time (setDiscretizationSamplingTime, NLMPC.hpp:80)
 *     __host__ __device__ static int nineq(int ph);            // Tineq
 *     __device__ static void f(double* out, const double* x, const double* u, int stage, const double* p);   // model
 *     __device__ static double cost(const Acc& a, double slack, int ph, const double* p);                    // objective
 *     __device__ static double ineq(int r, const Acc& a, double slack, int ph, const double* p);             // c_r <= 0
 *   optionally a sparsity hint  static constexpr int ineq_per_stage = K;   // inequality r reads only row r / K of (X, U)
 *   (rows that do not read the perturbed variable have an exactly-zero finite difference and are not evaluated),
 *   and optionally (user equality constraints)
 *     __host__ __device__ static int neq(int ph);   __device__ static double eq(int r, const Acc& a, int ph, const double* p);
 * where a.x(i,j) / a.u(i,j) read the unwrapped state / input sequences ((ph+1) rows, Mapping::unwrapVector).  An output
 * map (setOutputFunction, NLMPC.hpp:202) is  __device__ static void out(double* y, const double* x, const double* u, int i,
 * const double* p);  cost / ineq read it as  b200mpc::nl_y<Self>(a, i, j, p); b200mpc_nlmpc_output applies it to a solution
 * (OptSequence::output).  Returns a system id >= B200MPC_SYS_USER_BASE usable
 * wherever a built-in id is.  Kernels are compiled on first use and cached per device.  Compile errors are returned as
 * B200MPC_EINVAL with the NVRTC log in b200mpc_last_error(). */
#define B200MPC_SYS_USER_BASE 100
int b200mpc_nlmpc_register_system(const char* cuda_source, const char* type_name, int* system_id);
/* Compile-only check, needs no GPU: kernel 0 = evaluation kernel, 1..4 = the four solve-kernel variants, 5 = plant step / RK4. */
int b200mpc_nlmpc_compile_check(const char* cuda_source, const char* type_name, int kernel, size_t* cubin_bytes);

/* NLMPC::setStateScale / setInputScale (NLMPC.hpp:108-130 -> Mapping::setStateScaling / setInputScaling,
 * Mapping.hpp:108-150): host arrays [nx] / [nu] of positive factors, either may be NULL (= 1). */
typedef struct {
    const double* state_scale;
    const double* input_scale;
} b200mpc_nlmpc_scaling;

int b200mpc_nlmpc_eval(int system, int ph, int ch, int batch, const double* z, const double* x0, const double* params,
                       int params_per_instance, double* fval, double* grad, double* ceq, double* Jeq, double* cin,
                       double* Jin, int dev, void* stream);

/* ---- NLMPC solve -------------------------------------------------------------------------------
 * Replaces NLOptimizer::run (include/mpc/NLMPC/NLOptimizer.hpp:412-638: nlopt::opt(LD_SLSQP) with the objective,
 * the dynamics equalities, the user inequalities and the variable bounds) for `batch` controllers in one launch, one
 * warp per controller, shared-memory-resident damped-BFGS SQP (libmpc_b200/csrc/nlmpc_sqp.cuh).
 *   z0[batch*nz]: initial decision vectors (the reference's warm start: previous optimum shifted, NLOptimizer.hpp:430-470);
 *   lb/ub[nz]: variable bounds (Constraints.hpp bounds getters; +-inf for free variables; slack lower bound 0 when hard
 *   constraints are off, NLOptimizer.hpp:221-260); shared by the batch.
 *   outputs: z[batch*nz], cost[batch], viol[batch] (sum |c_eq| + sum max(c_in,0) at the solution),
 *   status[batch] (0 converged, 1 iteration limit), iters[batch] (major iterations), qp_iters[batch].
 * Problems whose matrices fit shared memory (b200mpc_nlmpc_solve_smem_bytes <= 227 KB) run entirely on chip; larger
 * ones keep the matrices in a per-warp HBM workspace (L2-resident) and only the vectors in shared memory. */
typedef struct {
    int32_t max_sqp;      /* major iterations (NLParameters::maximum_iteration, default 100)                 */
    int32_t max_qp;       /* ADMM iteration cap per QP subproblem (200; raised 5x once a line search fails)        */
    double tol;           /* relative step tolerance (NLParameters::relative_xtol plays this role)           */
    double ftol;          /* stop when |g'd| < ftol*max(1,|f|) and feasible (NLParameters::relative_ftol)   */
    double qp_eps;        /* QP residual tolerance                                                            */
    double rho;           /* initial ADMM penalty                                                             */
} b200mpc_nlmpc_params;
/* As b200mpc_nlmpc_eval, plus the scaling and the user equality constraints Constraints::evaluateEq
 * (NLMPC/Constraints.hpp:365-442,731-832): cue[batch*neq], Jue[batch*neq*nz] (ignored when the system has none). */
int b200mpc_nlmpc_eval_ex(int system, int ph, int ch, int batch, const double* z, const double* x0, const double* params,
                          int params_per_instance, const b200mpc_nlmpc_scaling* scaling, double* fval, double* grad,
                          double* ceq, double* Jeq, double* cin, double* Jin, double* cue, double* Jue, int dev,
                          void* stream);
/* OptSequence::output as NLOptimizer::run fills it (NLOptimizer.hpp:596-611): Model::getOutput (NLMPC/Model.hpp:72-96) applied
 * to every row of the sequences Mapping::unwrapVector gives for z.  y[batch*(ph+1)*ny]; all zeros for a system without an
 * output map, exactly as the reference (Model.hpp:82-84). */
int b200mpc_nlmpc_output(int system, int ph, int ch, int batch, const double* z, const double* x0, const double* params,
                         int params_per_instance, const b200mpc_nlmpc_scaling* scaling, double* y, int dev, void* stream);
void b200mpc_nlmpc_default_params(b200mpc_nlmpc_params* p);
/* Which solve kernel b200mpc_nlmpc_solve* / _closed_loop launch (process-wide).  0 (default): automatic -- the STAGE-STRUCTURED
 * kernel (libmpc_b200/csrc/nlmpc_structured.cuh: per-stage block BFGS, compact Jacobians, bordered block-tridiagonal KKT
 * factorisation, everything of a controller in shared memory) whenever the system declares `ineq_per_stage`, has no user equality
 * constraints and fits shared memory, else the dense kernel (nlmpc_sqp.cuh); 1: dense; 2: structured (EINVAL when it does not
 * apply).  Both reach the same optimum (tests/test_gpu_nlmpc_structured.py); the iteration paths differ (dense vs block BFGS). */
int b200mpc_nlmpc_set_solver(int solver);
long long b200mpc_nlmpc_solve_smem_bytes(int system, int ph, int ch);
int b200mpc_nlmpc_solve(int system, int ph, int ch, int batch, const b200mpc_nlmpc_params* params, const double* z0,
                        const double* x0, const double* sys_params, int params_per_instance, const double* lb,
                        const double* ub, double* z, double* cost, double* viol, int32_t* status, int32_t* iters,
                        int32_t* qp_iters, int dev, void* stream);

/* As b200mpc_nlmpc_solve with state / input scaling (the decision vector z is in scaled units, as in the reference;
 * user equality constraints of the system, if any, are always enforced). */
int b200mpc_nlmpc_solve_ex(int system, int ph, int ch, int batch, const b200mpc_nlmpc_params* params, const double* z0,
                           const double* x0, const double* sys_params, int params_per_instance,
                           const b200mpc_nlmpc_scaling* scaling, const double* lb, const double* ub, double* z,
                           double* cost, double* viol, int32_t* status, int32_t* iters, int32_t* qp_iters, int dev,
                           void* stream);

/* ---- set-up helper: mpc::RK4<N>::run(t, in, h, integration_step) (include/mpc/Integrator.hpp:16-56) for `batch`
 * states in one launch.  The vector field is the system's model with the input held over the step, dx/dt = f(x, u, stage, params);
 * as in the reference the time argument (`stage` here) is not advanced between the sub-steps.  x[batch*nx], u[batch*nu],
 * x_out[batch*nx]. */
int b200mpc_nlmpc_rk4(int system, int batch, int stage, const double* x, const double* u, const double* sys_params,
                      int params_per_instance, double h, int integration_steps, double* x_out, int dev, void* stream);

/* ---- NLMPC closed loop on the device: the loop the examples wrap around optimize()
 * (examples/vanderpol_ex.cpp:76-85, ugv_ex.cpp:143-166) for `steps` control steps without leaving the GPU.  Step k:
 *   1. NLOptimizer::run's initial guess (NLOptimizer.hpp:425-510) built by a kernel: cold tile of (x_k, u_{k-1}) on the first step
 *      or when enable_warm_start == 0, else the previous optimum; fixOptimalSolution (:705-716); one-stage left shift of the state
 *      rows and of the per-stage controls through Iz2u / Iu2z (move blocking); slack carry-over;
 *   2. the batched solve (as b200mpc_nlmpc_solve_ex);
 *   3. cmd = first control block (x input scaling) applied to the plant = the system's own model:
 *      plant_mode 0: x+ = f(x,u) (discrete systems); 1: x+ = x + plant_h f(x,u) (the Euler step of vanderpol_ex.cpp:80-81);
 *      2: plant_substeps RK4 steps of size plant_h (Integrator.hpp).
 * Outputs: traj_x[(steps+1)*batch*nx], traj_u[steps*batch*nu]; traj_status / traj_iters[steps*batch], traj_cost[steps*batch] may be
 * NULL.  No host round trip between the steps; synchronises at the end. */
int b200mpc_nlmpc_closed_loop(int system, int ph, int ch, int batch, const b200mpc_nlmpc_params* params, const double* x0,
                              const double* u0, const double* sys_params, int params_per_instance,
                              const b200mpc_nlmpc_scaling* scaling, const double* lb, const double* ub, int steps,
                              int enable_warm_start, int plant_mode, int plant_substeps, double plant_h, double* traj_x,
                              double* traj_u, int32_t* traj_status, int32_t* traj_iters, double* traj_cost, int dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MPC_H */
