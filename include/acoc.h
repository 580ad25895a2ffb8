/*
 * acoc.h -- C ABI of libacoc.so: batched regularized-Newton trajectory optimisation for the 2-D
 * longitudinal aircraft model on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the hot path of MohamedAtwan/AirCraftOptimalControl.  The reference
 * is pure Python with no FFI of its own, so each entry point below cites the Python interface it
 * replaces (file:line into the reference); INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add to route those functions here.
 *
 * Conventions
 *   - every function returns 0 on success and a negative acoc_status on failure; nothing throws or
 *     calls exit() across the boundary (the reference print()+exit()s on shape errors, optcon.py:585-596);
 *     acoc_last_error() returns a human-readable message for the calling thread.
 *   - all floating-point buffers are C-contiguous IEEE float64 HOST memory unless the name ends in _dev.
 *   - host trajectory layout is the reference's: one instance is (6,TT) / (2,TT) component-major
 *     (xx[i*TT + t]); a batch adds a leading instance axis (N,6,TT) / (N,2,TT).
 *   - matrices are row-major; "A" means fx.T (A[i][j] = d f_i / d x_j), "B" means fu.T.
 *   - params[9] = {cd0, cda, cla, m, g, S, rho, J, dt}          (aircraft_simplified.py:108-118)
 *   - there is no CPU fallback: every entry point runs CUDA kernels on the selected device and fails with
 *     ACOC_ERR_CUDA when no B200-class device is usable.
 */
#ifndef ACOC_H
#define ACOC_H

#ifdef __cplusplus
extern "C" {
#endif

#define ACOC_VERSION 100 /* 0.1.0 */

typedef enum acoc_status {
    ACOC_OK = 0,
    ACOC_ERR_INVALID = -1, /* bad argument (shape, NULL, option out of range) */
    ACOC_ERR_CUDA = -2,    /* CUDA runtime error, message in acoc_last_error() */
    ACOC_ERR_STATE = -3,   /* call sequence error (e.g. iterate before set_init) */
    ACOC_ERR_NOMEM = -4
} acoc_status;

/* flags for acoc_ctx_create */
#define ACOC_STATE_F32 0u        /* round the next state to float32 like aircraft_simplified.py:300 (default) */
#define ACOC_STATE_F64 1u        /* keep the next state in float64 (the reference with line 300 patched) */
#define ACOC_REFS_SHARED 2u      /* one reference trajectory shared by all instances */
#define ACOC_ARMIJO_SPECULATIVE 0u /* evaluate all armijo_maxiters candidates concurrently (default) */
#define ACOC_ARMIJO_LAZY 4u      /* candidate 0 for everyone, the remaining candidates only for instances that failed it */
#define ACOC_SOLVE_IN_PLACE 8u   /* acoc_newton_solve: never gather the still-iterating instances into a smaller survivor generation */
#define ACOC_FP32 16u            /* optional FP32 mode: float32 arithmetic and trajectory storage (costs/descent still accumulated in
                                    float64); NOT the parity path -- results agree with the float64 path to a tolerance only */
#define ACOC_NO_TMA 64u          /* run the sweeps with plain global loads instead of the warp-private TMA (bulk async copy) rings;
                                    results are bit-identical, this is for A/B measurements */
#define ACOC_NO_SPLIT 128u        /* never sweep a fully active batch as two tile ranges on two streams (A/B measurements; identical results) */
#define ACOC_NO_FUSED 256u       /* never fuse the LQ forward pass with line-search rollouts: neither with candidate 0 of the lazy search
                                    (k_forward_cand0_tma, any batch size) nor with the whole search of small batches (k_search_fused, <= 4096
                                    instances); the separate sweeps run instead (A/B tests; identical results) */
#define ACOC_PRIORITY_SHIFT 16
#define ACOC_PRIORITY(level) (((unsigned)(level) & 15u) << ACOC_PRIORITY_SHIFT)
                                 /* stream priority level of the context, 0 (default) .. 15: the device schedules the thread blocks of a
                                    context with a higher level first (clamped to the device's priority range).  Contexts that solve
                                    concurrently on one GPU (the sub-batches of a pipelined solve) then finish one after the other instead of
                                    all at the end, so that the device->host copy of one overlaps the iterations of the next.  No effect
                                    on results. */
#define ACOC_REFS_EXPANDED 512u  /* acoc_set_refs_generated: write the generated references out as per-instance arrays (64 B per
                                    instance and step, as acoc_set_refs would hold them) instead of keeping them in their parametric
                                    form (8 B or none); bit-identical results, A/B tests */
#define ACOC_X_F64 32u           /* keep the state iterates in float64 device buffers even when every stored state is a float32 value
                                    (ACOC_STATE_F32); results are bit-identical either way, this only costs bandwidth (A/B tests) */

/* per-instance status written by the Newton driver */
#define ACOC_INST_ACTIVE 0
#define ACOC_INST_CONVERGED 1   /* descent >= term_cond (optcon.py:499) */
#define ACOC_INST_MAXITER 2     /* ran max_iters-1 loop bodies (optcon.py:415) */
#define ACOC_INST_NONFINITE 3   /* NaN/Inf cost or descent: frozen */

/* acoc_newton_options.method: which optimize() loop acoc_newton_iterate / acoc_newton_solve run.
 *   ACOC_METHOD_NEWTON    NewtonMethod.optimize (optcon.py:341-529).
 *   ACOC_METHOD_GRADIENT  GradientMethod.optimize (optcon.py:27-174), steepest descent: deltau_t = -B_t' lam_{t+1} - lu_t (:111),
 *       descent = sum_t |deltau_t|^2 (:118), stop when descent <= -term_cond (1e-6 hard-coded at :52).  The reference's call of its own
 *       line search (optcon.py:125) passes 8 of armijo_stepsize's 9 arguments and raises TypeError; the repair used here (and by the
 *       oracle: oracle/pyref.py::run_gradient) supplies the missing JP = JJ[kk] and hands the search the directional derivative
 *       -descent, so that the test of optcon.py:268 is the sufficient-decrease condition.  Histories report the slope -descent. */
#define ACOC_METHOD_NEWTON 0
#define ACOC_METHOD_GRADIENT 1

typedef struct acoc_ctx acoc_ctx;

/* NewtonMethod.__init__ keyword arguments (optcon.py:335-337) */
typedef struct acoc_newton_options {
    int max_iters;        /* 200 in the shipped scripts (main_newton_method.py:32) */
    int armijo_maxiters;  /* 10   (main_newton_method.py:38) */
    int exact_after;      /* exact Hessian iff kk > exact_after; the reference hard-codes 8 (optcon.py:443) */
    int method;           /* ACOC_METHOD_NEWTON (0, default) or ACOC_METHOD_GRADIENT: which optimize() the driver runs (see below) */
    double stepsize_0;    /* 1    (main_newton_method.py:33) */
    double cc;            /* 0.5 */
    double beta;          /* 0.7 */
    double term_cond;     /* the reference hard-codes -1e-6 and ignores the constructor's value (optcon.py:368) */
} acoc_newton_options;

/* ---------------------------------------------------------------- library ---------------------------- */
int acoc_version(void);
const char* acoc_last_error(void);
int acoc_device_count(int* count);
/* name[len], SM count, memory bytes, compute capability major*10+minor of `device` */
int acoc_device_info(int device, char* name, int len, int* sm_count, unsigned long long* mem_bytes, int* cc);

/* ------------------------------------------------- pointwise entry points ---------------------------- */
/* Dynamics.step(xx, uu[, lmbd])  -- aircraft_simplified.py:263-393 (+ tensorCont :397-404)
 * n independent samples: x[n][6], u[n][2], lmbd[n][6] or NULL.
 * Outputs (any may be NULL): xxp[n][6]; A[n][36]; B[n][12] (6x2);
 *   lmbd == NULL: fxx[n][216] ([i][j][k] = d2 f_k/dx_i dx_j), fux[n][72] ([a][j][k]);
 *   lmbd != NULL: fxx[n][36], fux[n][12] contracted with the costate.   fuu is identically zero (:382). */
int acoc_step_batch(int device, int n, const double* params, int state_f64, const double* x, const double* u,
                    const double* lmbd, double* xxp, double* A, double* B, double* fxx, double* fux);

/* Cost.stagecost / Cost.termcost -- aircraft_simplified.py:25-69, :71-97.  n samples x[n][6], u[n][2],
 * xr[n][6], ur[n][2]; dense Q[36], R[4], QT[36].  Outputs (NULL to skip): ll[n], lx[n][6], lu[n][2],
 * llT[n], lTx[n][6].  The constant Hessians lxx=Q, luu=R, lxu=lux=0, lTxx=QT are the caller's inputs. */
int acoc_cost_batch(int device, int n, const double* Q, const double* R, const double* QT, const double* x,
                    const double* u, const double* xr, const double* ur, double* ll, double* lx, double* lu,
                    double* llT, double* lTx);

/* ltv_LQR(AAin,BBin,QQin,RRin,SSin,QQfin,TT,x0,qq,rr,qqf) -- optcon.py:533-771 (= lqr_tracking.py:6-242)
 * nb independent problems, time-major: A[nb][TT][6][6], B[nb][TT][6][2], Q[nb][TT][6][6], R[nb][TT][2][2],
 * S[nb][TT][2][6], Qf[nb][6][6], x0[nb][6]; affine terms q[nb][TT][6], r[nb][TT][2], qf[nb][6] or all three
 * NULL (non-augmented branch).  n = 7 when augmented else 6.
 * Outputs: K[nb][TT][2][n], P[nb][TT][n][n] (may be NULL), xout[nb][TT][6], uout[nb][TT][2],
 * n_reg[nb] (may be NULL) = number of steps that took the +0.5*I branch (optcon.py:745-749). */
int acoc_ltv_lqr(int device, int nb, int TT, const double* A, const double* B, const double* Q, const double* R,
                 const double* S, const double* Qf, const double* x0, const double* q, const double* r,
                 const double* qf, double* K, double* P, double* xout, double* uout, int* n_reg);

/* lqr_tracking(xx_opt, uu_opt, tt) -- lqr_tracking.py:245-283, batched over n perturbations delta[n][6]
 * of the initial state.  xx_opt (6,TT), uu_opt (2,TT).  Outputs xx_reg[n] (6,TT), uu_reg[n] (2,TT) and
 * (optional) the shared gains K[TT][2][6]. */
int acoc_lqr_tracking(int device, int n, int TT, const double* params, int state_f64, const double* Q,
                      const double* R, const double* QT, const double* xx_opt, const double* uu_opt,
                      const double* delta, double* xx_reg, double* uu_reg, double* K);

/* Device time (CUDA events, ms) of the kernels of this thread's last acoc_lqr_tracking call: ms[0] total, ms[1] linearisation along
 * the nominal + the shared LQ solve (k_step_batch + k_lq_dense<6>), ms[2] the closed-loop rollouts (k_track). */
int acoc_last_pointwise_timing(double* ms);

/* ------------------------------------------------- batched Newton context ---------------------------- */
/* One context = one GPU, one CUDA stream, N OCP instances with horizon TT held struct-of-arrays in HBM.
 * A context is not thread-safe; different contexts are independent. */
int acoc_ctx_create(int device, int n_instances, int TT, unsigned flags, acoc_ctx** out);
int acoc_ctx_destroy(acoc_ctx* ctx);
/* bytes of device memory owned by the context */
int acoc_ctx_device_bytes(const acoc_ctx* ctx, unsigned long long* bytes);

int acoc_set_model(acoc_ctx* ctx, const double* params);                                  /* Dynamics() attrs */
/* Cost(QQt,RRt,QQT), aircraft_simplified.py:20-23.  Q, R and QT must be SYMMETRIC here (ACOC_ERR_INVALID otherwise): the fused
 * backward sweep carries the Riccati matrix as its upper triangle.  This is a documented deviation: the reference accepts any matrix
 * (aircraft_simplified.py:61-68) -- a non-symmetric Q there makes the gradient Q dx inconsistent with the cost dx'Q dx, which no
 * shipped configuration does.  The pointwise acoc_cost_batch and the literal acoc_ltv_lqr accept arbitrary matrices like the
 * reference (tests: test_weight_symmetry_contract). */
int acoc_set_weights(acoc_ctx* ctx, const double* Q, const double* R, const double* QT);
int acoc_set_options(acoc_ctx* ctx, const acoc_newton_options* opt);                      /* NewtonMethod(...) */
/* xx_ref (N,6,TT) / uu_ref (N,2,TT), or (6,TT) / (2,TT) when the context was created with ACOC_REFS_SHARED */
int acoc_set_refs(acoc_ctx* ctx, const double* xx_ref, const double* uu_ref);
/* The reference generators of the scripts, on the device (main_newton_method.py:96-142, acrobatic_newton.py:99-154): per-instance
 * xx_ref / uu_ref are built in HBM from per-instance parameters zf[N] (final / bump height), vx[N] (= (xf - x0)/tf) and shared time
 * bases tt[TT], zshape[TT], vshape[TT] (NULL: V_ref = xconst[2]) that the caller computed with the scripts' own numpy code:
 *   X = 0 + vx*tt,  Z = 0 + zshape*(zf - 0),  V = ((vshape*zf)**2 + vx**2)**0.5,  theta, q, gamma = xconst[3..5],  uu_ref = uconst[0..1]
 * in the scripts' operation order, so the references are bit-identical to what the scripts build on the host and acoc_set_refs
 * would upload (16 B per instance cross the bus instead of 64 KB).  The context keeps them in this parametric form -- the sweeps form
 * X and Z from the tables with one multiplication each, only V (step maneuver) is a stored per-instance array -- which removes 56-64 of
 * the ~230-280 bytes a sweep moves per instance and step (ACOC_REFS_EXPANDED keeps per-instance arrays instead; same results up to the
 * sign of a zero reference).  Not for ACOC_REFS_SHARED contexts. */
int acoc_set_refs_generated(acoc_ctx* ctx, const double* tt, const double* zshape, const double* vshape, const double* zf,
                            const double* vx, const double* xconst, const double* uconst);
/* the references held by the context, in acoc_set_refs' layout (either pointer may be NULL) */
int acoc_get_refs(acoc_ctx* ctx, double* xx_ref, double* uu_ref);
/* xx_init (N,6,TT), uu_init (N,2,TT): NewtonMethod.optimize(xx_init, uu_init, tf, dt), optcon.py:341, :395-398.
 * Resets the Newton state (iteration counter, histories). */
int acoc_set_init(acoc_ctx* ctx, const double* xx_init, const double* uu_init);
/* Dynamics.get_initial_trajectory(xx_ref, tt) -- aircraft_simplified.py:126-148 -- computed on the device for
 * every instance from the references already set (float64 arithmetic), then used as the initial guess.
 * dx0 (N,6) or NULL: start instance i from xx_ref[:,0] + dx0[i] (perturbed-initial-state batches). */
int acoc_init_guess(acoc_ctx* ctx, double kp, double kt, const double* dx0);

/* Run up to n_iters more Newton iterations (loop bodies of optcon.py:415-501) on every instance that is
 * still active; returns after the device has finished them (the call is timed with CUDA events, see
 * acoc_get_timing).  *n_active_out (may be NULL) = instances still active afterwards. */
int acoc_newton_iterate(acoc_ctx* ctx, int n_iters, int* n_active_out);
/* Iterate until no instance is active.  *total_iters = sum over instances of loop bodies executed.
 * Whenever at most half of a batch (of >= 4096 instances) is still iterating, the survivors are gathered into a smaller
 * internal context so that every warp lane does useful work again; they are folded back before the call returns, so
 * results, histories and statistics are exactly those of iterating in place (ACOC_SOLVE_IN_PLACE disables this). */
int acoc_newton_solve(acoc_ctx* ctx, long long* total_iters);
/* Block until all work queued on the context's stream has finished. */
/* acoc_newton_solve and the read-back of what optimize() returns in ONE call.  xx_star: (N,6,TT) float32 if x_is_f32 (lossless under the
 * float32 state quantisation, see acoc_get_result_f32; column 0 then holds float32(x0)) else float64; uu_star (N,2,TT) float64; x0 (N,6)
 * or NULL receives the exact initial states.  When xx_star and uu_star are PAGE-LOCKED host memory (cudaHostAlloc / cudaHostRegister,
 * e.g. torch pin_memory) the layout-conversion kernels write straight into them, and the instances that have finished are delivered at
 * the moment the still-iterating ones move on to a survivor generation, so the transfer overlaps the latency-bound tail of the solve.
 * Pageable buffers: same results through acoc_newton_solve + acoc_get_result[_f32]. */
int acoc_newton_solve_deliver(acoc_ctx* ctx, void* xx_star, int x_is_f32, double* uu_star, double* x0, long long* total_iters);
int acoc_sync(acoc_ctx* ctx);

/* What NewtonMethod.optimize returns (optcon.py:503-505): iterate kk-1 with uu[:, -1] = uu[:, -2]. */
int acoc_get_result(acoc_ctx* ctx, double* xx_star, double* uu_star);
/* The same result with the states as float32 (N,6,TT): LOSSLESS whenever the context keeps its state iterates as float (float32
 * state quantisation of aircraft_simplified.py:300, the default): every x_t, t >= 1, is a float32 value.  Column t = 0 holds
 * float32(x0); the exact float64 x0 = xx_init[:,0] (optcon.py:398) is returned in x0 (N,6) when non-NULL.  uu_star stays float64.
 * ACOC_ERR_STATE if the context holds float64 states (ACOC_STATE_F64, ACOC_X_F64, or acoc_set_init with non-float32 states). */
int acoc_get_result_f32(acoc_ctx* ctx, float* xx_star, double* uu_star, double* x0);
/* which: 0 = the newest iterate (last get_update), 1 = the one before it. */
int acoc_get_iterate(acoc_ctx* ctx, int which, double* xx, double* uu);
/* Descent direction of the last iteration: deltau (N,2,TT) (optcon.py:468).  deltau and the gains are per-iteration
 * scratch: they are meaningful for instances that were still active in that iteration. */
int acoc_get_deltau(acoc_ctx* ctx, double* deltau);
/* Gains of the last backward sweep: K (N,2,6,TT), sigma (N,2,TT) -- KK[:,1:,:] and KK[:,0,:] of optcon.py:468.
 * Stages N*16*TT doubles on the host: meant for inspection / parity tests on small batches. */
int acoc_get_gains(acoc_ctx* ctx, double* K, double* sigma);
/* Per-instance, per-iteration history, each (N, max_iters): JJ[k], descent[k] (optcon.py:497), the Armijo step
 * (optcon.py:327) and the number of candidates the sequential search rolls out.  Any pointer may be NULL. */
int acoc_get_history(acoc_ctx* ctx, double* JJ, double* descent, double* stepsize, int* n_cand);
/* Per-instance summary (length N each, any may be NULL): loop bodies executed, ACOC_INST_* status, cost of
 * the newest iterate, last descent, number of +0.5*I regularisations. */
int acoc_get_stats(acoc_ctx* ctx, int* iters, int* status, double* J, double* descent, int* n_reg);

/* Single pieces of the loop on the context's current iterate, exposed for parity tests and for the
 * reference's own method names:
 *   acoc_eval_cost        optcon.py:417-424   -> J[N]
 *   acoc_backward         optcon.py:429-464 + Riccati/gains of ltv_LQR; exact != 0 uses the exact Hessian (:443)
 *   acoc_forward          optcon.py:756-762 + :474-477 -> descent[N]
 *   acoc_armijo           GradientMethod.armijo_stepsize (optcon.py:204-327) -> stepsize[N], costs[N][armijo_maxiters]
 *   acoc_update           GradientMethod.get_update (optcon.py:176-200) with per-instance steps -> becomes the newest iterate */
int acoc_eval_cost(acoc_ctx* ctx, double* J);
/* Overwrite the descent direction deltau (N,2,TT) and/or the scalars the Armijo test uses (JP = J[N],
 * descent[N]); lets GradientMethod.armijo_stepsize / get_update be called with caller-supplied arguments. */
int acoc_set_deltau(acoc_ctx* ctx, const double* deltau);
int acoc_set_scalars(acoc_ctx* ctx, const double* J, const double* descent);
int acoc_backward(acoc_ctx* ctx, int exact);
int acoc_forward(acoc_ctx* ctx, double* descent);
int acoc_armijo(acoc_ctx* ctx, double* stepsize, double* costs);
/* One costate sweep of GradientMethod.optimize (optcon.py:95-118) on the current iterate: deltau (readable with acoc_get_deltau)
 * and descent[N] = sum_t |deltau_t|^2 (may be NULL).  The context keeps -descent as the slope of the following acoc_armijo. */
int acoc_gradient(acoc_ctx* ctx, double* descent);
/* The visu_armijo sweep (optcon.py:280-296): costs[N][n_steps] of the rollouts u + steps[k]*deltau for caller-chosen step sizes
 * (the reference plots np.linspace(0, stepsize_0, 10)) along the current deltau. */
int acoc_armijo_sweep(acoc_ctx* ctx, int n_steps, const double* steps, double* costs);
int acoc_update(acoc_ctx* ctx, const double* stepsize);

/* Device-side timing of the last acoc_newton_iterate call, in milliseconds (CUDA events on the context's
 * stream): total and per phase {cost, backward, forward, candidates, select, update}; plus launch count. */
int acoc_get_timing(acoc_ctx* ctx, double* total_ms, double phase_ms[6], long long* kernel_launches);
/* Enable (1) / disable (0) per-phase event timing; off by default because it adds events to the stream. */
int acoc_set_profiling(acoc_ctx* ctx, int on);

/* FP64 FMA throughput microbenchmark (register-resident DFMA chains on every SM): the measured denominator
 * for the FP64 side of the roofline.  Returns TFLOP/s (2 flops per DFMA). */
int acoc_measure_fp64_peak(int device, double* tflops);
/* Cycles per operation of one warp's chain of dependent DFMAs (8.7 on B200): the in-order latency a lone warp of the
 * sequential sweeps pays per dependent instruction. */
int acoc_measure_fp64_latency(int device, double* cycles_per_op);
/* Device copy bandwidth microbenchmark (read + write bytes / time), GB/s. */
int acoc_measure_copy_bw(int device, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* ACOC_H */
