/*
 * slode_b200 -- C ABI of the B200-native latent-ODE solve (sm_100a).
 *
 * Drop-in boundary for the ONE hot path of paidamoyo/structured_latent_ODEs: the batched
 * integration of the latent ODE and its reverse-mode gradient.  The reference has no FFI of its
 * own (it is pure Python); the boundary it crosses is the third-party call
 *
 *     torchdiffeq.odeint_adjoint(func=, y0=, t=, method=)   models/blackbox_ode.py:41-42
 *     torchdiffeq.odeint(func=, y0=, t=, method=)           models/blackbox_ode.py:44-45
 *
 * with func = OdeFunc(z, Dynamics) (models/blackbox_ode.py:50-61, 64-109).  Each entry point
 * below replaces one leg of that call for one right-hand side.  The Python host
 * (structured_latent_odes_b200/torchdiffeq_api.py) binds these symbols with ctypes inside a
 * torch.autograd.Function; INTEGRATION.md shows the two-line change on the reference side.
 *
 * Conventions (all entry points)
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch); the library never frees
 *     or retains it; outputs are caller-allocated;
 *   - all arrays are float32; `stream` is a cudaStream_t passed as void*; work is enqueued
 *     asynchronously on it and the call returns immediately;
 *   - return value 0 = ok, otherwise an SLODE_E* code; slode_last_error() gives the message of
 *     the last failure on the calling thread;
 *   - the fixed-grid entry points keep no state between calls: weights are staged per block from the
 *     caller's tensors, scratch memory is the caller's (`workspace`, sized by slode_fixed_workspace_bytes),
 *     so they are re-entrant, may run concurrently on any streams and can be captured into CUDA graphs.
 *     (The dopri5 entry points still share one library-owned scratch area per device and are serialised.)
 *
 * The blackbox right-hand side (Dynamics.forward, models/blackbox_ode.py:97-109):
 *     x = [t, z];  h = relu(W1 x + b1);  f = sigmoid(Wg h + bg) - sigmoid(Wd h + bd) * state
 * enters as   h_j(t) = relu(w1t[j] * t + c[b][j])   with
 *     w1t = W1[:, 0]            (the column that multiplies t, :72,:101)
 *     c   = z @ W1[:, 1:]^T + b1   (B,H)  -- time-invariant per trajectory; computed inside the kernels
 *                                            by the slode_latent_* entry points (from z), or handed in
 *                                            precomputed to the slode_mlp_* entry points
 * Wg/Wd are dyanamics_growth.weight / dyanmics_degradation.weight, torch Linear layout (S,H).
 */
#ifndef SLODE_B200_H
#define SLODE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes */
#define SLODE_OK 0
#define SLODE_EINVAL 1      /* bad argument (null pointer, negative size, unknown method/mode) */
#define SLODE_EUNSUPPORTED 2 /* (H,S) / option combination not compiled in: no generic fallback */
#define SLODE_ECUDA 3       /* a CUDA runtime call failed; see slode_last_error() */

/* torchdiffeq `method=` strings the reference can pass (config.solver, data/x/config_x.py) */
#define SLODE_METHOD_EULER 0
#define SLODE_METHOD_MIDPOINT 1
#define SLODE_METHOD_RK4 2      /* torchdiffeq's rk4 == 3/8 rule (rk4_alt_step_func) */
#define SLODE_METHOD_DOPRI5 3

/* backward modes */
#define SLODE_BWD_DISCRETE 0 /* exact gradient of the unrolled solver == torchdiffeq.odeint + autograd
                                (adjoint_solver=False, models/blackbox_ode.py:44-45) */
#define SLODE_BWD_TDE_ADJOINT 1 /* torchdiffeq.odeint_adjoint semantics (adjoint_solver=True, :41-42):
                                   continuous adjoint re-discretised with the same method, state reset
                                   to the stored forward value at every output time */

/* slode_query(what) */
#define SLODE_Q_VERSION 0
#define SLODE_Q_SM_ARCH 1        /* 100 */
#define SLODE_Q_MAX_HIDDEN 2     /* largest H compiled for the register-resident kernels */
#define SLODE_Q_MAX_STATE 3      /* largest S */
#define SLODE_Q_N_SHAPES 4       /* number of compiled (H,S) pairs */
#define SLODE_Q_SHAPE_BASE 100   /* 100+2i -> H of pair i, 101+2i -> S of pair i */
#define SLODE_Q_FWD_LAUNCHES 10  /* kernels launched by the last forward entry-point call of the process */
#define SLODE_Q_BWD_LAUNCHES 11
#define SLODE_Q_TOTAL_LAUNCHES 12 /* kernels launched by this library since it was loaded (all threads) */
#define SLODE_Q_SOURCE_HASH 13    /* 31-bit digest of the sources this binary was built from (_build.source_hash) */

int slode_query(int what);
const char* slode_last_error(void);

/* 1 if fixed-grid kernels (slode_*_fixed_*) for this (H,S) pair exist, else 0 */
int slode_mlp_supported(int H, int S);
/* 1 if the dopri5 kernels (slode_mlp_dopri5_*) exist for this (H,S) pair, else 0 */
int slode_dopri5_supported(int H, int S);

/*
 * Scratch memory of the fixed-grid entry points below, in bytes, for a call with these arguments (0 is
 * possible).  The caller allocates it on the device (16-byte aligned: anything cudaMalloc / the torch allocator gives),
 * passes it as `workspace` and may free or reuse it once the call's work on `stream` has finished.  Contents
 * need not be initialised and are not preserved.
 *   backward     0: slode_*_fixed_fwd, 1: slode_*_fixed_bwd (flip records of the reverse sweep: NQ*16 bytes
 *                per hidden unit and RESIDENT thread -- it does not grow with B beyond one wave of blocks)
 *   fused        0: slode_mlp_fixed_* (c, y0 given), 1: slode_latent_* with y0 given, 2: with the x0 net fused
 *   rows_in_time 1 if sol is (B,T,S)-contiguous (sol_stride_t == S)
 * Returns -1 for arguments the entry point itself would reject.
 */
int64_t slode_fixed_workspace_bytes(int backward, int method, int mode, int64_t B, int T, int L, int H, int S,
                                    int fused, int rows_in_time);

/*
 * Forward fixed-grid solve; replaces torchdiffeq.odeint(func=OdeFunc, y0, t, method) for
 * method in {euler, midpoint, rk4} with grid == t (models/blackbox_ode.py:44-45).
 *   t    (T)    strictly monotone output times == solver grid (torchdiffeq asserts the same).  The kernels follow
 *               every hidden unit's ReLU crossing along the sweep (the heads are piecewise linear in t along a
 *               trajectory), which is only defined for a monotone grid; the Python entry points check it, a host
 *               that binds this ABI directly must guarantee it -- a non-monotone t gives wrong results, not an error.
 *   c    (B,H)  row-major, see above;   y0 (B,S) row-major
 *   sol  out: element (i,b,s) at sol[i*sol_stride_t + b*sol_stride_b + s]; sol[0] = y0.
 *        (T,B,S)-contiguous (torchdiffeq's layout) is stride_t=B*S, stride_b=S;
 *        (B,T,S)-contiguous (what solve_ODE's permute hands the decoder) is stride_t=S, stride_b=T*S.
 */
int slode_mlp_fixed_fwd(int method, int64_t B, int T, int H, int S,
                        const float* t, const float* c, const float* y0,
                        const float* w1t, const float* Wg, const float* bg,
                        const float* Wd, const float* bd,
                        float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                        void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Reverse-mode gradient of slode_mlp_fixed_fwd (one reverse sweep over the stored grid states;
 * stages are recomputed from sol[i], nothing else is checkpointed).
 *   grad_sol  upstream dL/dsol, same indexing convention as sol (own strides)
 *   grad_y0   out (B,S);   grad_c  out (B,H)  (host turns it into dz, dW1[:,1:], db1 with cuBLAS)
 *   grad_w    in/out, ACCUMULATED into (caller zero-fills): flat
 *             [ dw1t (H) | dWg (S*H) | dbg (S) | dWd (S*H) | dbd (S) ]
 * mode SLODE_BWD_DISCRETE    : exact discrete adjoint (parity with odeint + autograd);
 * mode SLODE_BWD_TDE_ADJOINT : odeint_adjoint emulation (grad_c then only feeds dW1/db1: the
 *                              reference drops dz through the dynamics in this mode).
 */
int slode_mlp_fixed_bwd(int method, int mode, int64_t B, int T, int H, int S,
                        const float* t, const float* c,
                        const float* w1t, const float* Wg, const float* bg,
                        const float* Wd, const float* bd,
                        const float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                        const float* grad_sol, int64_t gsol_stride_t, int64_t gsol_stride_b,
                        float* grad_y0, float* grad_c, float* grad_w,
                        void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Fused variants of slode_mlp_fixed_fwd / _bwd: the solve together with the two small nets in front of it, i.e.
 * the whole of OdeModel.solve_ODE (models/blackbox_ode.py:36-47):
 *     c  = z W1[:,1:]^T + b1                               (first layer of Dynamics on the constants, :99-106)
 *     x0 = sigmoid(Wb relu(Wa z + ba) + bb)                (latent_to_ode_net, :19-22, :32-34)
 * computed per trajectory inside the solver kernels (no (B,H) intermediates in HBM, no cuBLAS calls).  The reverse
 * sweep re-evaluates the heads from sol[i] (piecewise-linear evaluator); nothing but sol is checkpointed.
 *   z (B,L);  W1 = dynamics_hidden.weight (H, L+1), b1 its bias;  Wa (H,L), ba (H), Wb (S,H), bb (S) =
 *   latent_to_ode_net[0] / [2] -- pass all four as NULL and give y0 (B,S) instead when the caller computes the
 *   initial state itself (the torchdiffeq.odeint entry, where y0 is an argument).
 * Backward: grad_z (B,L) written (through c -- discrete mode only, SURVEY F5 -- and through x0 when the x0 net is
 * fused); grad_y0 (B,S) written when the x0 net is NOT fused (else NULL); grad_params ACCUMULATED (caller
 * zero-fills), flat
 *     [ dw1t (H) | dWg (S*H) | dbg (S) | dWd (S*H) | dbd (S) | dW1[:,1:] (H*L) | db1 (H) |
 *       dWa (H*L) | dba (H) | dWb (S*H) | dbb (S) ]           (the last four only when the x0 net is fused).
 */
int slode_latent_fixed_fwd(int method, int64_t B, int T, int L, int H, int S,
                           const float* t, const float* z,
                           const float* W1, const float* b1, const float* Wg, const float* bg,
                           const float* Wd, const float* bd,
                           const float* Wa, const float* ba, const float* Wb, const float* bb,
                           const float* y0,
                           float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                           void* workspace, int64_t workspace_bytes, void* stream);

/*
 * The forward solve with the decoder heads fused into its epilogue; replaces, in ONE kernel,
 *     solution = OdeModel.solve_ODE(z)                       models/blackbox_ode.py:36-47
 *     mu_q     = Linear_q(solution).permute(0, 2, 1)         models/decoders.py:45-47 (Decoder), :86 (GaussianDecoder)
 * for callers that use the head outputs only (reconstruction / posterior sampling under no_grad,
 * training_challenge.py:174-195, training_proc.py:205-223): the heads are applied to the states while the solver
 * thread still holds them, mu is written in the reference's (B, obs_dim, T) layout, and the trajectories themselves
 * reach HBM only if the caller asks for them.
 *   head_W  (NQ, O, S) contiguous, as in slode_heads_fwd
 *   mu      out: element (q, b, o, t) at mu[((q*B + b)*O + o)*mu_row_pitch + t], mu_row_pitch >= T floats.
 *           mu_row_pitch = T is the reference's contiguous (B, obs_dim, T) per head.  The kernel writes whole 32-byte
 *           sectors of a row (eight consecutive times) per store; with a pitch that is a multiple of 8 floats (T
 *           rounded up; the caller hands out mu[..., :T] views) every row has the same sector phase and all
 *           trajectories of a warp store at the same steps -- about a fifth faster than an unaligned pitch.
 *   sol     NULL (not written) or as in slode_latent_fixed_fwd with its strides
 *   every other argument as in slode_latent_fixed_fwd; workspace: slode_fixed_workspace_bytes of the forward with
 *   the same sizes (rows_in_time = 0) is sufficient.  Results are bit-equal to slode_latent_fixed_fwd followed by
 *   slode_heads_fwd.  Forward only: a training step needs sol as the reverse sweep's checkpoint and uses the two
 *   separate entry points.
 */
int slode_latent_fixed_heads_fwd(int method, int64_t B, int T, int L, int H, int S,
                                 const float* t, const float* z,
                                 const float* W1, const float* b1, const float* Wg, const float* bg,
                                 const float* Wd, const float* bd,
                                 const float* Wa, const float* ba, const float* Wb, const float* bb,
                                 const float* y0,
                                 int O, int NQ, const float* head_W, float* mu, int64_t mu_row_pitch,
                                 float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                                 void* workspace, int64_t workspace_bytes, void* stream);

int slode_latent_fixed_bwd(int method, int mode, int64_t B, int T, int L, int H, int S,
                           const float* t, const float* z,
                           const float* W1, const float* b1, const float* Wg, const float* bg,
                           const float* Wd, const float* bd,
                           const float* Wa, const float* ba, const float* Wb, const float* bb,
                           const float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                           const float* grad_sol, int64_t gsol_stride_t, int64_t gsol_stride_b,
                           float* grad_z, float* grad_y0, float* grad_params,
                           void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Adaptive Dormand-Prince 5(4) forward solve; replaces torchdiffeq.odeint(func=OdeFunc, y0, t, method="dopri5",
 * rtol, atol) (models/blackbox_ode.py:44-45 with config.solver = "dopri5").  torchdiffeq semantics are kept:
 * ONE step size for the whole batch (error ratio = RMS over all B*S elements), float64 controller time,
 * Hairer initial step, FSAL, 4th-order dense output at the requested times.  The whole adaptive loop runs in one
 * persistent cooperative kernel; the accepted-step sequence is deterministic.
 *   t            (T) strictly monotone output times (float32).  Decreasing times run torchdiffeq's reversed solve
 *                (_ReverseFunc: s = -t, negated right-hand side) as the same scheme with negative steps; step_log
 *                and replay_steps then hold the caller's t0 and dt < 0
 *   first_step   > 0 to skip the initial-step selection (torchdiffeq options["first_step"]), else <= 0
 *   max_attempts bound on attempted steps (torchdiffeq options["max_num_steps"])
 *   replay_steps optional (n_replay, 3) float64 in the step_log format: take exactly these step sizes and
 *                accept/reject decisions instead of running the controller (re-running a logged solve; parity
 *                tests against a reference step sequence); NULL for the normal adaptive solve
 *   ckpt_y       optional (ckpt_capacity, B, S): state at the start of every accepted step (the checkpoints the
 *                reverse sweep needs); NULL to skip
 *   step_log     optional (log_capacity, 3) float64: t0, dt, accepted(1/0) of every attempted step; NULL to skip
 *   stats        device int64[4]: accepted steps, rejected steps, RHS evaluations per trajectory, status
 *                (0 ok, 1 dt underflow, 2 max_attempts exceeded, 3 ckpt_capacity exceeded, 4 replay too short)
 */
int slode_mlp_dopri5_fwd(int64_t B, int T, int H, int S,
                         const float* t, const float* c, const float* y0,
                         const float* w1t, const float* Wg, const float* bg,
                         const float* Wd, const float* bd,
                         double rtol, double atol, double first_step, int64_t max_attempts,
                         const double* replay_steps, int64_t n_replay,
                         float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                         float* ckpt_y, int64_t ckpt_capacity,
                         double* step_log, int64_t log_capacity,
                         int64_t* stats, void* stream);

/*
 * The same solve advanced ONE PASS PER CALL, for a batch that is sharded over several devices / processes
 * (SURVEY.md section 8(e)): torchdiffeq's controller is batch-global, so every shard has to see the error norm of the
 * WHOLE batch.  Each call runs one pass over this shard (Hairer's two initial-step passes, then one attempted step
 * per call), writes the shard's sums of squares to out_sums[2] and returns; the host adds out_sums over the shards
 * (NCCL all-reduce) into ext_sums[2] and calls again.  All shards then take the accept / reject decisions and step
 * sizes of the unsharded solve.  Controller state and trajectory state live in the caller's workspace between calls.
 *   n_global   trajectories over all shards (the norms divide by n_global * S)
 *   restart    1 on the first call of a solve, 0 afterwards
 *   stats      device int64[5]: as for slode_mlp_dopri5_fwd, plus [4] = phase after this call (5 = solve finished;
 *              stop calling).  A shard may not be empty.  replay_steps are not supported here.
 *   workspace  slode_mlp_dopri5_step_workspace_bytes(B, S) bytes, 256-byte aligned, caller-owned, kept between calls
 */
int64_t slode_mlp_dopri5_step_workspace_bytes(int64_t B, int S);

int slode_mlp_dopri5_fwd_step(int64_t B, int T, int H, int S,
                              const float* t, const float* c, const float* y0,
                              const float* w1t, const float* Wg, const float* bg,
                              const float* Wd, const float* bd,
                              double rtol, double atol, double first_step, int64_t max_attempts,
                              int64_t n_global, int restart,
                              const double* ext_sums, double* out_sums,
                              float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                              float* ckpt_y, int64_t ckpt_capacity,
                              double* step_log, int64_t log_capacity,
                              int64_t* stats, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Reverse-mode gradient of slode_mlp_dopri5_fwd: exact gradient of the accepted-step sequence (what autograd
 * through torchdiffeq.odeint(method="dopri5") gives; step sizes and accept/reject decisions carry no gradient).
 *   accepted_steps (n_accepted, 2) float64: t0, dt of every accepted step (rows of step_log with accepted = 1)
 *   emit_ranges    (n_accepted + 1) int32: output times [emit[n], emit[n+1]) were interpolated in accepted step n
 *                  (emit[0] = 1: sol[0] is y0 itself)
 *   ckpt_y         the checkpoints written by the forward call
 *   grad_* as for slode_mlp_fixed_bwd (grad_w accumulated into; caller zero-fills).
 * (The torchdiffeq.odeint_adjoint variant of dopri5 is slode_mlp_dopri5_adjoint_bwd below.)
 */
int slode_mlp_dopri5_bwd(int64_t B, int T, int H, int S,
                         const float* t, const float* c,
                         const float* w1t, const float* Wg, const float* bg,
                         const float* Wd, const float* bd,
                         int64_t n_accepted, const double* accepted_steps, const int* emit_ranges,
                         const float* ckpt_y,
                         const float* grad_sol, int64_t gsol_stride_t, int64_t gsol_stride_b,
                         float* grad_y0, float* grad_c, float* grad_w,
                         void* stream);

/*
 * Backward pass of torchdiffeq.odeint_adjoint(func=OdeFunc, y0, t, method="dopri5", rtol, atol)
 * (models/blackbox_ode.py:40-42 with config.solver = "dopri5"; adjoint_solver=True is the shipped default of every
 * config).  torchdiffeq semantics are kept: for i = T-1 .. 1 a FRESH adaptive dopri5 solve of the augmented system
 * [y, a, a_theta] from t[i] down to t[i-1] (time flipped, Hairer initial step per interval, one step size for the
 * whole augmented state, mixed error norm = the largest per-tensor RMS among y, a and each parameter tensor's
 * adjoint, float64 controller time, FSAL, the interval's end value from the 4th-order interpolant), then y is
 * reset to the stored forward value sol[i-1] and a += grad_sol[i-1].  Gradients go to y0 and to
 * func.parameters() = (dynamics_hidden, dyanamics_growth, dyanmics_degradation) only: OdeFunc.constants (z) gets
 * none.  One persistent cooperative kernel; the step sequence is deterministic.
 *   t (T) increasing output times; z (B,L); c (B,H) = z W1[:,1:]^T + b1
 *   W1 (H, L+1) dynamics_hidden.weight (column 0 multiplies t); Wg,Wd (S,H); bg,bd (S)
 *   sol / grad_sol indexing as in slode_mlp_fixed_bwd (sol = the forward solve's output)
 *   grad_y0 (B,S) written; grad_params written, flat in func.parameters() order:
 *       [ dW1 (H*(L+1)) | db1 (H) | dWg (S*H) | dbg (S) | dWd (S*H) | dbd (S) ]
 *   step_log optional (log_capacity, 4) float64: interval index i, -t at the start of the attempt, step size,
 *       accepted(1/0), one row per attempted step; NULL to skip
 *   replay_steps optional (n_replay, 4) float64 rows in the step_log format: take exactly these step sizes and
 *       accept/reject decisions (parity tests against a reference step sequence); NULL for the adaptive solve
 *   stats    device int64[4]: accepted, rejected, RHS evaluations per trajectory, status (0 ok, 1 dt underflow,
 *       2 max_attempts exceeded, 4 replay too short)
 *   workspace: slode_mlp_dopri5_adjoint_workspace_bytes(B, L, H, S) bytes, 256-byte aligned, caller-owned
 * ode_state_dim in {4, 5, 8}; hidden width limited by shared memory (about 150 for L = 15).
 */
int64_t slode_mlp_dopri5_adjoint_workspace_bytes(int64_t B, int L, int H, int S);

int slode_mlp_dopri5_adjoint_bwd(int64_t B, int T, int L, int H, int S,
                                 const float* t, const float* z, const float* c,
                                 const float* W1, const float* Wg, const float* bg,
                                 const float* Wd, const float* bd,
                                 const float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                                 const float* grad_sol, int64_t gsol_stride_t, int64_t gsol_stride_b,
                                 double rtol, double atol, int64_t max_attempts,
                                 const double* replay_steps, int64_t n_replay,
                                 float* grad_y0, float* grad_params,
                                 double* step_log, int64_t log_capacity, int64_t* stats,
                                 void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Decoder heads on the latent trajectories; replaces, for all heads at once,
 *     mu_q = Linear_q(solution).permute(0, 2, 1)            models/decoders.py:45-47 (Decoder: q50, q75, q25)
 *     mean = output_mean(solution).permute(0, 2, 1)         models/decoders.py:86     (GaussianDecoder)
 *   sol   element (t,b,s) at sol[t*sol_stride_t + b*sol_stride_b + s]
 *   W     (NQ, O, S) contiguous: the NQ bias-free Linear(S -> O) weights stacked
 *   mu    out, (NQ, B, O, T) contiguous: mu[q] is head q in the reference's (B, obs_dim, T) layout
 * Backward: grad_mu (NQ,B,O,T) -> grad_sol (written, own strides) and grad_W (NQ,O,S) ACCUMULATED (caller
 * zero-fills).  ode_state_dim in {4, 5, 8}, obs_dim <= 8, NQ <= 3.
 */
int slode_heads_fwd(int64_t B, int T, int S, int O, int NQ,
                    const float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                    const float* W, float* mu, void* stream);

int slode_heads_bwd(int64_t B, int T, int S, int O, int NQ,
                    const float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                    const float* W, const float* grad_mu,
                    float* grad_sol, int64_t gsol_stride_t, int64_t gsol_stride_b,
                    float* grad_W, void* stream);

/* The same with one (B,O,T) gradient per head (autograd hands the quantile heads' gradients over separately;
 * NULL = that head did not enter the loss): no stacking copy in front of the kernel. */
int slode_heads_bwd_split(int64_t B, int T, int S, int O, int NQ,
                          const float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                          const float* W, const float* grad_mu0, const float* grad_mu1, const float* grad_mu2,
                          float* grad_sol, int64_t gsol_stride_t, int64_t gsol_stride_b,
                          float* grad_W, void* stream);

/* element types of the CVS entry points */
#define SLODE_F32 0
#define SLODE_F64 1

/*
 * CVS mechanistic right-hand side (dx_dt, data/cvs/cvs_data.py:52-91), 4 states (Pa/100, Pv/10, S, SV/100),
 * fixed-grid solve.  Replaces the reference's one-trajectory-at-a-time scipy LSODA loop
 * (create_cvs_data, data/cvs/cvs_data.py:111-134) with a batched solve, and gives the RHS the odeint-style
 * forward(t, state) module API of the latent ODE (north_star: "mechanistic RHS fused into the stage evaluation").
 *   dtype      SLODE_F32 (odeint drop-in) or SLODE_F64 (data generator: the reference generates in float64);
 *              every array below has that element type
 *   substeps   solver steps per output interval (1 = torchdiffeq's "grid == t"; <= 16)
 *   y0 (B,4), i_ext (B), r_tpr_mod (B): per-trajectory initial state and treatments (:24-26, :106-108)
 *   theta (10) = [f_hr_max, f_hr_min, r_tpr_max, r_tpr_min, sv_mod, ca, cv, k_width, p_aset, tau] (:28-49)
 *   sol / grad_sol indexing as in slode_mlp_fixed_fwd with S = 4
 * Backward outputs: grad_y0 (B,4), grad_i_ext (B), grad_r_tpr_mod (B) written; grad_theta (10) ACCUMULATED
 * (caller zero-fills).  Modes as for slode_mlp_fixed_bwd.
 */
int slode_cvs_fixed_fwd(int method, int dtype, int64_t B, int T, int substeps,
                        const void* t, const void* y0, const void* i_ext, const void* r_tpr_mod,
                        const void* theta,
                        void* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                        void* stream);

int slode_cvs_fixed_bwd(int method, int mode, int dtype, int64_t B, int T, int substeps,
                        const void* t, const void* i_ext, const void* r_tpr_mod, const void* theta,
                        const void* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                        const void* grad_sol, int64_t gsol_stride_t, int64_t gsol_stride_b,
                        void* grad_y0, void* grad_i_ext, void* grad_r_tpr_mod, void* grad_theta,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLODE_B200_H */
