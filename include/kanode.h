/*
 * kanode.h — C ABI of the B200-native KAN-ODE hot path (libkanode_b200.so).
 *
 * This is the drop-in boundary described in SURVEY.md §8(b).  The reference
 * (maharshi-coding/KAN-ODEs) has no FFI: its "API" is three Julia call shapes.
 * Each entry point below names the reference call site it replaces
 * (LV/ = Lotka-Volterra/, PDE/ = "PDE examples/"):
 *
 *   kanode_create / kanode_set_params
 *       <- KDense(I,O,G; normalizer, basis_func, use_base_act)   LV/src/kdense.jl:20-68
 *          Lux.Chain(...), Lux.setup(rng, chain)                  LV/LV_driver_KANODE.jl:139-143
 *          flat parameter vector p (ComponentArray data)          LV/LV_driver_KANODE.jl:173-175
 *   kanode_rhs
 *       <- (l::KDense)(x, p, st)                                  LV/src/kdense.jl:109-130
 *          kan1(u, p, st) as ODE right-hand side                  LV/LV_driver_KANODE.jl:180
 *          rc_kanode(u, p, t) (Laplacian + pointwise KAN)         PDE/Allen-Cahn_Source.jl:90-93
 *   kanode_vjp
 *       <- Zygote.pullback of the chain (rrule(_rbf))             LV/src/utils.jl:15-21
 *   kanode_solve
 *       <- NeuralODE(kan, tspan, Tsit5(); saveat)(u0, p, st)      LV/LV_driver_KANODE.jl:180-184
 *          solve(ODEProblem(rc_kanode,u0,tspan,p;saveat),Tsit5()) PDE/Allen-Cahn_Source.jl:96-99
 *   kanode_loss_grad
 *       <- Zygote.gradient(loss, p)[1] with
 *          loss(p) = mean(abs2, X .- predict(p))                  LV/LV_driver_KANODE.jl:197-203,284
 *                                                                 PDE/Burgers_Surrogate.jl:105-107,191
 *
 * Conventions
 *  - All functions return 0 on success, a negative kanode_status on failure;
 *    kanode_last_error() gives the message.  No C++ exception crosses the ABI.
 *  - Per-trajectory solver outcomes (MaxIters, DtLessThanMin, Unstable) are NOT
 *    call failures: they are reported in kanode_stats.retcode.
 *  - Host entry points take host pointers and copy in/out on the handle's
 *    stream; *_dev entry points take device pointers (same layouts) and only
 *    enqueue work — call kanode_sync() (or synchronise the stream) afterwards.
 *  - There is NO CPU fallback: every compute entry point fails with
 *    KANODE_ERR_NO_DEVICE if no sm_100-class CUDA device is usable.
 *  - Layouts (column-major in Julia terms == "trajectory-major" here):
 *      u0     [batch][n]            one trajectory's state is contiguous
 *      out    [batch][nsave][n]     == Array(sol) (n x nsave) per trajectory
 *      target [batch][nsave][n]
 *      p      flat: per layer  C[O][G*I] column-major (column = i*G + g, i.e.
 *             element (o,i,g) at  (i*G+g)*O + o ), then W[O][I] column-major
 *             (element (o,i) at i*O + o)  — LV/src/kdense.jl:75,81 and
 *             LV/Activation_getter.jl:9-10.
 *  - Arithmetic type of the device path is fp32 for states/parameters and
 *    fp64 for the time variable; the reference runs Float64.
 */
#ifndef KANODE_H_
#define KANODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KANODE_MAX_LAYERS 8

typedef enum {
    KANODE_OK = 0,
    KANODE_ERR_INVALID = -1,      /* bad descriptor / argument                 */
    KANODE_ERR_NO_DEVICE = -2,    /* no usable CUDA device (no CPU fallback)   */
    KANODE_ERR_CUDA = -3,         /* CUDA runtime error, see kanode_last_error */
    KANODE_ERR_NOMEM = -4,
    KANODE_ERR_UNSUPPORTED = -5,  /* valid request outside the built kernels   */
    KANODE_ERR_SOLVER = -6        /* kanode_loss_grad*: some trajectory's solve did not return Success; outputs are written,
                                     the failed trajectories are left out of loss / grad (retcodes in the stats arrays) */
} kanode_status;

/* normalizer: LV/src/kdense.jl:25,41-47,57-61 (NNlib.fast_act maps tanh->tanh_fast) */
typedef enum {
    KANODE_NORM_TANH = 0,         /* tanh / tanh_fast  (LV_driver_KANODE.jl:131) */
    KANODE_NORM_SOFTSIGN = 1,     /* softsign          (Burgers_Surrogate.jl:83) */
    KANODE_NORM_SIGMOID = 2       /* sigmoid / sigmoid_fast                      */
} kanode_normalizer;

/* basis_func: LV/src/utils.jl:8-62 */
typedef enum {
    KANODE_BASIS_RBF = 0,         /* exp(-((x-z)/h)^2)   utils.jl:8-13  */
    KANODE_BASIS_RSWAF = 1,       /* 1 - tanh(((x-z)/h))^2  utils.jl:27-34 */
    KANODE_BASIS_IQF = 2          /* 1/(1+((x-z)/h)^2)   utils.jl:49-54 */
} kanode_basis;

/* rhs_kind */
typedef enum {
    KANODE_RHS_CHAIN = 0,         /* du = chain(u)                NeuralODE dudt       */
    KANODE_RHS_SOURCE_LAPLACIAN = 1, /* du = lap_coef*lap(u) + chain_1to1.(u)  AC_Source:90-93 */
    KANODE_RHS_MAP = 2            /* not an ODE: the chain as a map x[I_first] -> y[O_last], i.e. the direct layer call
                                     (l::KDense)(x, p, st) -> (y, st) of kdense.jl:109-130 (Activation_getter.jl:39,
                                     Allen-Cahn_Source.jl:91).  n_state = I_first.  kanode_rhs evaluates it (output
                                     [batch][O_last]), kanode_vjp pulls back; the solve entry points refuse it. */
} kanode_rhs_kind;

typedef struct {
    int32_t in_dims;              /* I */
    int32_t out_dims;             /* O */
    int32_t grid_len;             /* G */
    int32_t normalizer;           /* kanode_normalizer */
    int32_t basis;                /* kanode_basis      */
    int32_t use_base_act;         /* 1: y += W*swish(x)  (kdense.jl:122-127) */
    float   grid_lo;              /* grid_lims[1], default -1f0 (kdense.jl:26) */
    float   grid_hi;              /* grid_lims[2], default  1f0               */
    float   denominator;          /* h, default Float32(2/(G-1)) (kdense.jl:27) */
    int32_t kind;                 /* kanode_layer_kind: 0 = KDense (all fields above), 1 = Lux.Dense (MLP-NODE baseline) */
    int32_t dense_act;            /* kind 1: output activation (enum kanode_dense_act); identity on the last layer */
} kanode_layer_desc;

/* layer kind.  KANODE_LAYER_DENSE is `Lux.Dense(in => out, act)`: y = act(W x + b), the right-hand side of the MLP-NODE baseline
 * `Lux.Chain(Lux.Dense(2 => 50, tanh), Lux.Dense(50 => 2))` (Lotka-Volterra/LV_driver_MLP.jl:61).  Flat parameters of the layer in the
 * ComponentArray order of that driver (:65-67): [vec(weight[out, in]); bias[out]].  grid_len / normalizer / basis / use_base_act
 * are ignored for it.  Dense and KDense layers may be mixed in one chain; such chains run on the block-per-trajectory kernels. */
typedef enum { KANODE_LAYER_KDENSE = 0, KANODE_LAYER_DENSE = 1 } kanode_layer_kind;
typedef enum { KANODE_ACT_IDENTITY = 0, KANODE_ACT_TANH = 1 } kanode_dense_act;

typedef struct {
    int32_t n_layers;
    kanode_layer_desc layers[KANODE_MAX_LAYERS];
    int32_t rhs_kind;             /* kanode_rhs_kind */
    int32_t n_state;              /* n: state length of one trajectory */
    double  lap_coef;             /* s*D: AC_Source -1e-4, Fisher-KPP +0.01 (rhs_kind 1) */
    double  dx;                   /* grid spacing of the periodic 3-point Laplacian      */
} kanode_desc;

/* retcode mirrors SciML ReturnCode names */
typedef enum {
    KANODE_RET_SUCCESS = 0,
    KANODE_RET_MAXITERS = 1,
    KANODE_RET_DT_LESS_THAN_MIN = 2,
    KANODE_RET_UNSTABLE = 3,
    KANODE_RET_RECORD_OVERFLOW = 4 /* internal: dense-record capacity hit; host retries larger */
} kanode_retcode;

typedef struct {
    int32_t naccept;
    int32_t nreject;
    int32_t nf;                   /* RHS evaluations (backward: fused forward+VJP evals) */
    int32_t retcode;              /* kanode_retcode */
} kanode_stats;

typedef struct kanode_handle kanode_handle;

const char* kanode_version(void);
/* message of the last failure on this handle (or of the last failed create if h==NULL) */
const char* kanode_last_error(const kanode_handle* h);

/* number of parameters implied by a descriptor (LuxCore.parameterlength, kdense.jl:98-107);
 * returns 0 for an invalid descriptor. */
size_t kanode_param_count(const kanode_desc* desc);

/* device: CUDA device ordinal; stream: a cudaStream_t cast to void* (NULL = library-owned stream). */
int kanode_create(const kanode_desc* desc, int device, void* stream, kanode_handle** out);
int kanode_destroy(kanode_handle* h);
int kanode_sync(kanode_handle* h);

/* copy the flat parameter vector (np floats) to the device */
int kanode_set_params(kanode_handle* h, const float* p, size_t np);
int kanode_set_params_dev(kanode_handle* h, const float* d_p, size_t np);

/* du[b] = f(u[b]) for b < batch */
int kanode_rhs(kanode_handle* h, const float* u, float* du, int64_t batch);
int kanode_rhs_dev(kanode_handle* h, const float* d_u, float* d_du, int64_t batch);

/* ubar[b] = (df/du)^T lam[b];  pbar = sum_b (df/dp)^T lam[b]   (pbar has np floats) */
int kanode_vjp(kanode_handle* h, const float* u, const float* lam,
               float* ubar, float* pbar, int64_t batch);

/* adaptive Tsit5, defaults of the reference are abstol=1e-6, reltol=1e-3 */
int kanode_solve(kanode_handle* h, const float* u0, int64_t batch,
                 double t0, double t1, const double* saveat, int32_t nsave,
                 float abstol, float reltol,
                 float* out, kanode_stats* stats /* [batch] or NULL */);
int kanode_solve_dev(kanode_handle* h, const float* d_u0, int64_t batch,
                     double t0, double t1, const double* saveat /* host */, int32_t nsave,
                     float abstol, float reltol,
                     float* d_out, kanode_stats* d_stats /* device, [batch] or NULL */);

/* loss = mean over (batch, nsave, n) of (out - target)^2;  grad = dloss/dp (np floats).
 * Each trajectory runs the reference's interpolating-adjoint backward solve on
 * z=[lambda; g] with its own step control; grad = (1/batch) * sum_b g_b(t0).
 * du0 (optional, [batch][n]) receives dloss_b/du0_b (lambda_b(t0), unscaled by 1/batch). */
int kanode_loss_grad(kanode_handle* h, const float* u0, int64_t batch,
                     double t0, double t1, const double* saveat, int32_t nsave,
                     const float* target, float abstol, float reltol,
                     float* loss, float* grad, float* du0 /* or NULL */,
                     kanode_stats* fwd_stats /* [batch] or NULL */,
                     kanode_stats* bwd_stats /* [batch] or NULL */);
/* device-pointer variant.  d_loss_sum receives sum_b sum_{i,j}(out-target)^2 (a double, NOT yet
 * divided) and d_grad_sum receives sum_b g_b(t0) (np floats, NOT yet divided by batch), so a
 * data-parallel caller can all-reduce both and divide by the global counts. */
int kanode_loss_grad_dev(kanode_handle* h, const float* d_u0, int64_t batch,
                         double t0, double t1, const double* saveat /* host */, int32_t nsave,
                         const float* d_target, float abstol, float reltol,
                         double* d_loss_sum, float* d_grad_sum, float* d_du0 /* or NULL */,
                         kanode_stats* d_fwd_stats, kanode_stats* d_bwd_stats);

/* ---- Float64 entry points ------------------------------------------------------------------------
 * The reference drivers effectively run in Float64 (LV_driver_KANODE.jl:117,175 promote u0 and p), so a
 * Julia binding passes Vector{Float64}.  These run the same kernels instantiated for double: they reproduce
 * the reference's step sequence (same accepted-step count) where fp32 state arithmetic cannot (DESIGN.md,
 * "precision").  Same layouts and semantics as the float versions above. */
int kanode_set_params_f64(kanode_handle* h, const double* p, size_t np);
int kanode_rhs_f64(kanode_handle* h, const double* u, double* du, int64_t batch);
int kanode_vjp_f64(kanode_handle* h, const double* u, const double* lam,
                   double* ubar, double* pbar, int64_t batch);
int kanode_solve_f64(kanode_handle* h, const double* u0, int64_t batch,
                     double t0, double t1, const double* saveat, int32_t nsave,
                     double abstol, double reltol, double* out, kanode_stats* stats);
int kanode_loss_grad_f64(kanode_handle* h, const double* u0, int64_t batch,
                         double t0, double t1, const double* saveat, int32_t nsave,
                         const double* target, double abstol, double reltol,
                         double* loss, double* grad, double* du0,
                         kanode_stats* fwd_stats, kanode_stats* bwd_stats);
int kanode_loss_grad_dev_f64(kanode_handle* h, const double* d_u0, int64_t batch,
                             double t0, double t1, const double* saveat /* host */, int32_t nsave,
                             const double* d_target, double abstol, double reltol,
                             double* d_loss_sum, double* d_grad_sum, double* d_du0,
                             kanode_stats* d_fwd_stats, kanode_stats* d_bwd_stats);

/* ---- dt-replay (parity tooling; SURVEY.md §7.3 "replay reference dt sequence") ------------------------------------
 * Same result as kanode_loss_grad, but the step-size controller of BOTH solves is bypassed: the forward solve takes the
 * accepted steps whose END times are fwd_t[b][0..] (ascending, NaN-padded to max_steps), the adjoint solve those of
 * bwd_t[b][0..] (descending, the save times among them).  Feeding the fp64 oracle's sequences to the fp32 kernels separates
 * arithmetic parity from controller parity: what [EXT OrdinaryDiffEqCore] decides (a12-a14 of SURVEY.md §8a) is taken from
 * the reference run, what remains is the kernels' arithmetic.  out (optional) receives the predictions [batch][nsave][n].
 * Only the small-model ensemble kernels implement it (KANODE_ERR_UNSUPPORTED otherwise). */
int kanode_loss_grad_replay(kanode_handle* h, const float* u0, int64_t batch,
                            double t0, double t1, const double* saveat, int32_t nsave,
                            const float* target, float abstol, float reltol,
                            const double* fwd_t, const double* bwd_t, int32_t max_steps,
                            float* loss, float* grad, float* du0 /* or NULL */, float* out /* or NULL */,
                            kanode_stats* fwd_stats /* or NULL */, kanode_stats* bwd_stats /* or NULL */);
int kanode_loss_grad_replay_f64(kanode_handle* h, const double* u0, int64_t batch,
                                double t0, double t1, const double* saveat, int32_t nsave,
                                const double* target, double abstol, double reltol,
                                const double* fwd_t, const double* bwd_t, int32_t max_steps,
                                double* loss, double* grad, double* du0, double* out,
                                kanode_stats* fwd_stats, kanode_stats* bwd_stats);

/* ---- pullback of the solve for an ARBITRARY loss (the reference's gradient call, LV_driver_KANODE.jl:197-203,284 with the
 * optional reg term; Burgers_Surrogate.jl:105-107,191 with a transposed target) -------------------------------------
 * Zygote.gradient(loss, p) runs the forward solve, differentiates loss(pred) itself and hands dL/dpred to the adjoint of
 * the solve [EXT SciMLSensitivity 7.69.0 InterpolatingAdjoint].  This entry point is that adjoint: the forward problem is
 * re-solved densely, the backward problem on z = [lambda; g] is integrated t1 -> t0 with the jumps lambda += dL_dout[b][s]
 * at the save times.  grad = sum_b (d pred_b/d p)^T dL_dout[b] (no 1/batch: the cotangent carries every scale),
 * du0[b] = (d pred_b/d u0_b)^T dL_dout[b].  out (optional) receives the predictions of the dense forward solve.
 * Returns KANODE_ERR_SOLVER like kanode_loss_grad when a solve fails.  The Julia rrule on the NeuralODE call binds this. */
int kanode_solve_adjoint(kanode_handle* h, const float* u0, int64_t batch, double t0, double t1,
                         const double* saveat, int32_t nsave, float abstol, float reltol,
                         const float* dL_dout /* [batch][nsave][n] */, float* out /* or NULL */,
                         float* grad /* [np] */, float* du0 /* [batch][n] or NULL */,
                         kanode_stats* fwd_stats /* or NULL */, kanode_stats* bwd_stats /* or NULL */);
int kanode_solve_adjoint_f64(kanode_handle* h, const double* u0, int64_t batch, double t0, double t1,
                             const double* saveat, int32_t nsave, double abstol, double reltol,
                             const double* dL_dout, double* out, double* grad, double* du0,
                             kanode_stats* fwd_stats, kanode_stats* bwd_stats);
/* device-pointer variant (no host synchronisation; d_grad receives the same sum, d_out optional) */
int kanode_solve_adjoint_dev(kanode_handle* h, const float* d_u0, int64_t batch, double t0, double t1,
                             const double* saveat, int32_t nsave, float abstol, float reltol,
                             const float* d_dL_dout, float* d_out, float* d_grad, float* d_du0,
                             kanode_stats* d_fwd_stats, kanode_stats* d_bwd_stats);

/* ---- per-edge activations (LV/Activation_getter.jl:3-63; feeds prune, LV_driver_KANODE.jl:52-108, and the plotters) ----
 * act[k][i][o] = sum_g C_l[o,(i,g)] * basis_g(normalizer(x[k][i])) + W_l[o,i] * swish(x[k][i]) for layer `layer` of the chain
 * at its inputs x[K][I_l]: the fused basis kernel without the sum over the inputs; sum_i act[k][i][o] is the layer output
 * (the identity commented at Activation_getter.jl:33-36). */
int kanode_edge_activations(kanode_handle* h, int32_t layer, const float* x /* [K][I_l] */, float* act /* [K][I_l][O_l] */, int64_t K);
int kanode_edge_activations_f64(kanode_handle* h, int32_t layer, const double* x, double* act, int64_t K);

/* ---- sparsity regulariser (reg_loss, LV_driver_KANODE.jl:187-194; added to the loss when sparse_on == 1, :199-201) ----
 * reg(p) = act_reg * sum|p| + entropy_reg * (-sum e log e), e = |p| / sum|p|.  Once set (non-zero), kanode_loss_grad* add
 * reg(p) to the loss and d reg/d p to the gradient (the *_dev variants add batch*nsave*n*reg and batch*dreg to their
 * un-normalised sums, so the caller's normalisation gives the same).  (0, 0) switches it off (the default). */
int kanode_set_regularizer(kanode_handle* h, double act_reg, double entropy_reg);
/* reg(p) and d reg/d p of the current parameters on their own (host pointers; grad may be NULL) */
int kanode_reg_loss(kanode_handle* h, double act_reg, double entropy_reg, double* loss, float* grad /* [np] or NULL */);

/* ---- several GPUs behind ONE handle (the Julia surface stays single-process) ---------------------------
 * kanode_create_multi builds one child handle per listed device (own stream, own workspace) inside the returned handle.  The
 * host-pointer entry points kanode_set_params*, kanode_solve*, kanode_loss_grad*, kanode_solve_adjoint*,
 * kanode_set_regularizer and kanode_set_record_capacity accept it: the batch is split into contiguous shards, one host thread
 * per device drives its shard (trajectories are independent: no data-path exchange), and the only cross-device step is the sum
 * of the per-device gradient / loss sums, done by ONE kernel on the first device that loads the other devices' partial sums
 * straight from their memory over NVLink peer access (staged with cudaMemcpyPeer when peer access is unavailable); the sum
 * order is fixed, so the result does not depend on timing.  kanode_rhs* / kanode_vjp* / kanode_edge_activations* run on the
 * first device; the *_dev, *_replay and kanode_train_* entry points take device pointers of one GPU and return
 * KANODE_ERR_UNSUPPORTED on a multi-device handle (data-parallel processes use one plain handle each + kanode_train_apply_dev). */
int kanode_create_multi(const kanode_desc* desc, const int32_t* devices, int32_t n_devices, kanode_handle** out);
int32_t kanode_device_count(const kanode_handle* h);   /* 1 for a plain handle */

/* ---- device-resident training iteration (LV_driver_KANODE.jl:280-291: grad = Zgrad(loss, p)[1]; update!(opt, p, grad);
 * loss_train(p); loss_test(p)) --------------------------------------------------------------------------------------
 * kanode_train_begin copies the handle's current parameters into a device-resident fp32 master copy and zeroes the Adam
 * moments (Flux.Adam(eta, (beta1, beta2), eps), :219).  kanode_train_step_dev then runs ONE reference iteration without
 * any host synchronisation on the small-model ensemble path (the lockstep PDE engines poll a device flag while they
 * enqueue step attempts): loss + gradient at p_k (regulariser included when set), Adam update p_k -> p_{k+1}, refresh of
 * every derived device image of the parameters by kernels, then the two forward-only loss solves of the reference at
 * p_{k+1}.  d_losses (device, 3 doubles) receives loss(p_k), loss_train(p_{k+1}) = mean(abs2, target - predict) and
 * loss_test(p_{k+1}) (the test solve is skipped when d_target_test is NULL; its slot is left untouched).
 * All data pointers are device pointers of this handle's GPU; saveat arrays are host pointers.  fp32 handles only.
 * kanode_train_apply_dev is the second half on its own (Adam + refresh) for callers that all-reduce d_grad_sum across
 * processes first (g = grad_scale * d_grad_sum).  kanode_train_params copies the master parameters to the host (blocks). */
int kanode_train_begin(kanode_handle* h, float eta, float beta1, float beta2, float eps);
int kanode_train_step_dev(kanode_handle* h, const float* d_u0, int64_t batch, double t0, double t1,
                          const double* saveat, int32_t nsave, const float* d_target, float abstol, float reltol,
                          const float* d_u0_test /* or NULL */, int64_t batch_test, double t1_test,
                          const double* saveat_test, int32_t nsave_test, const float* d_target_test /* or NULL */,
                          double* d_losses /* device [3] */);
int kanode_train_apply_dev(kanode_handle* h, const float* d_grad_sum, float grad_scale);
int kanode_train_params(kanode_handle* h, float* p /* [np] host */);

/* ---- data-parallel plumbing (one process per GPU): ONE collective per step ---------------------------------------------
 * kanode_pack_sums_dev packs the un-normalised sums of kanode_loss_grad_dev* into one fp64 buffer
 *   d_packed[np + 2] = [gradient sum (np) | loss sum | trajectory count]
 * so that the caller all-reduces (NCCL over NVLink) exactly one buffer per step; kanode_train_apply_packed_dev then applies
 * Adam with g = packed gradient / packed count (read on the device: no host round trip) and refreshes the parameter images. */
int kanode_pack_sums_dev(kanode_handle* h, const float* d_grad_sum, const double* d_loss_sum, int64_t count, double* d_packed);
int kanode_pack_sums_dev_f64(kanode_handle* h, const double* d_grad_sum, const double* d_loss_sum, int64_t count, double* d_packed);
int kanode_train_apply_packed_dev(kanode_handle* h, const double* d_packed);

/* ---- the step's collective fused into the packing kernel, over NVLink peer memory (one process per GPU, one node) -------
 * The reference has no multi-device path; the data-parallel step of SURVEY.md 8(e) is "gradient sum -> all-reduce".  For the
 * small models ([gradient | loss | count] <= KANODE_PEER_MAX_ENTRIES doubles) that collective is latency, not bytes, so the
 * library does it itself instead of calling NCCL: every process owns a MAILBOX in its GPU's memory, exported to its peers as a
 * CUDA IPC handle.  kanode_pack_allreduce_dev launches ONE kernel that packs the sums, stores them straight into every peer's
 * mailbox over NVLink, publishes an epoch flag (release, system scope), waits for the peers' flags and adds the world's
 * contributions in rank order — every rank gets bit-identical sums, nothing goes through the host.
 *   kanode_peer_export   allocates the mailbox and writes its 64-byte cudaIpcMemHandle_t to `ipc_handle`
 *   kanode_peer_attach   opens the mailboxes of all ranks (`ipc_handles` = world x 64 bytes, in rank order; exchange them with
 *                        any host-side all-gather); world <= KANODE_PEER_MAX_WORLD; world == 1 needs no peers
 *   kanode_pack_allreduce_dev(_f64)   d_packed[np + 2] = sum over ranks of [gradient sum | loss sum | count]; every rank must
 *                        make the same sequence of calls.  A peer that does not arrive within ~20 s poisons the result with
 *                        NaN and the next kanode_peer_status returns KANODE_ERR_SOLVER instead of hanging the GPU.
 *   kanode_peer_status   blocks until the handle's stream is idle; 0, or the error of a timed-out exchange */
#define KANODE_PEER_MAX_WORLD 16
#define KANODE_PEER_MAX_ENTRIES 4096
#define KANODE_IPC_HANDLE_BYTES 64
int kanode_peer_export(kanode_handle* h, void* ipc_handle /* [64] host */);
int kanode_peer_attach(kanode_handle* h, int32_t rank, int32_t world, const void* ipc_handles /* [world][64] host */);
int kanode_pack_allreduce_dev(kanode_handle* h, const float* d_grad_sum, const double* d_loss_sum, int64_t count, double* d_packed);
int kanode_pack_allreduce_dev_f64(kanode_handle* h, const double* d_grad_sum, const double* d_loss_sum, int64_t count, double* d_packed);
int kanode_peer_status(kanode_handle* h);

/* Flux.Adam(eta, (beta1, beta2), eps) + update!(opt, p, grad)  (LV_driver_KANODE.jl:219,287; [EXT Flux 0.14.22]):
 *   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g^2;  p -= eta * (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps),  g = grad_scale*d_grad.
 * All pointers are device pointers of np floats; t is the 1-based iteration count.  grad_scale lets a data-parallel
 * caller pass the all-reduced gradient SUM and divide by the global trajectory count in the same kernel.
 * The handle's parameters are NOT changed: call kanode_set_params_dev(h, d_p, np) afterwards. */
int kanode_adam_step_dev(kanode_handle* h, float* d_p, const float* d_grad, float* d_m, float* d_v, int64_t t,
                         float eta, float beta1, float beta2, float eps, float grad_scale);

/* dense-record capacity (accepted forward steps kept per trajectory for the adjoint).  Host entry points grow
 * it automatically on overflow; *_dev entry points report KANODE_RET_RECORD_OVERFLOW in the stats instead. */
int kanode_set_record_capacity(kanode_handle* h, int32_t max_steps);

/* device time of the kernels of the LAST kanode_loss_grad*() call, measured with CUDA events recorded on the
 * handle's stream around each launch: ms[0] = forward (Tsit5 + dense record + loss), ms[1] = backward (adjoint),
 * ms[2] = gradient reduction.  Blocks until that call has finished. */
int kanode_last_timing(kanode_handle* h, float* ms3);

/* wide (batched lockstep) engine only: device time [ms] spent in the step-end passes over the per-IC gradient state g
 * (the HBM-bound kernels wide_gp1/wide_gp2) during the LAST kanode_loss_grad*() call, and the number of such passes
 * (= lockstep step attempts).  CUDA events around each pass on the handle's stream.  Fails if the last call did not
 * go through the wide engine. */
int kanode_last_gpass_timing(kanode_handle* h, float* ms, int32_t* passes);

/* number of kernel launches issued by this handle since creation (for bench accounting) */
int64_t kanode_launch_count(const kanode_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* KANODE_H_ */
