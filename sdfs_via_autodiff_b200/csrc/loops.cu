// Host entry points of the device-resident solver loops (kernels: loops.cuh; one translation
// unit per operator type instantiates them).
#include "loops.cuh"

static int build_env(sdfs_op *op, LoopEnv *env) {
    sdfs_ctx *ctx = op->ctx;
    memset(env, 0, sizeof(*env));
    env->rank = ctx->rank;
    env->nranks = ctx->nranks;
    env->status = (LoopStatus *)ctx->d_status;
    const int64_t N = op_N(op);
    const bool sharded = op_is_sharded(op);     // row-sharded dense P or slab-sharded factor form
    if (!sharded) {   // single GPU, or an operator that is whole on this rank: purely local loop
        env->rank = 0;
        env->nranks = 1;
        if (!op->slots) {
            CUDA_TRY(ctx, cudaMalloc(&op->slots, arena_slots_doubles() * sizeof(double)));
            CUDA_TRY(ctx, cudaMemsetAsync(op->slots, 0, arena_slots_doubles() * sizeof(double), ctx->stream));
        }
        env->slots[0] = op->slots;
        env->xin[0][0] = op->work;
        env->xin[0][1] = op->work + op->ldv;
        env->flags[0] = nullptr;
        env->epoch0 = 0;
    } else {
        if (!comm_peers_ready(ctx))
            return sdfs_set_error(ctx, SDFS_ERR_COMM, "multi-GPU solver loops need the exchange arena (sdfs_comm_arena_export/import)");
        if (comm_arena_maxN(ctx) < N)
            return sdfs_set_error(ctx, SDFS_ERR_ARG, "exchange arena sized for N<=%lld, operator has N=%lld", (long long)comm_arena_maxN(ctx), (long long)N);
        for (int r = 0; r < ctx->nranks; ++r)
            arena_carve(comm_peer_arena(ctx, r), comm_arena_maxN(ctx), &env->flags[r], &env->slots[r], &env->xin[r][0], &env->xin[r][1]);
        env->epoch0 = *comm_epoch(ctx);
    }
    return SDFS_OK;
}

static int finish_loop(sdfs_ctx *ctx, LoopStatus *hs, unsigned long long epochs_hint) {
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(hs, ctx->d_status, sizeof(LoopStatus), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    (void)epochs_hint;
    if (hs->abort_code != 0)
        return sdfs_set_error(ctx, SDFS_ERR_TIMEOUT, "device loop aborted: a peer rank did not reach the barrier");
    static const bool trace = getenv("SDFS_LOOP_TRACE") && atoi(getenv("SDFS_LOOP_TRACE")) != 0;
    if (trace && hs->n_apply)
        fprintf(stderr, "loop trace: %llu factor-form applications, %.1f us each inside the operator (modes %.1f %.1f %.1f %.1f, "
                        "epilogue phase %.1f), loop total %.3f ms\n", hs->n_apply, hs->t_apply_ns * 1e-3 / hs->n_apply,
                hs->t_mode_ns[0] * 1e-3 / hs->n_apply, hs->t_mode_ns[1] * 1e-3 / hs->n_apply, hs->t_mode_ns[2] * 1e-3 / hs->n_apply,
                hs->t_mode_ns[3] * 1e-3 / hs->n_apply, hs->t_epi_ns * 1e-3 / hs->n_apply, hs->t_total_ns * 1e-6);
    if (trace && hs->n_apply && hs->t_vec_ns[0])
        fprintf(stderr, "loop trace: BiCGSTAB vector phases per application: C %.1f us, E (+reduction) %.1f, G (+reduction) %.1f, reduction after D %.1f\n",
                hs->t_vec_ns[0] * 1e-3 / hs->n_apply, hs->t_vec_ns[1] * 1e-3 / hs->n_apply, hs->t_vec_ns[2] * 1e-3 / hs->n_apply,
                hs->t_red_ns * 1e-3 / hs->n_apply);
    return SDFS_OK;
}

int loop_launch_kron_sa(sdfs_op *op, void *a, LoopEnv *env);
int loop_launch_kron_newton(sdfs_op *op, void *a, LoopEnv *env);
int loop_launch_kron_anderson(sdfs_op *op, void *a, LoopEnv *env);
int loop_launch_kron(sdfs_op *op, int which, void *a, LoopEnv *env) {
    if (which == LOOP_SA) return loop_launch_kron_sa(op, a, env);
    if (which == LOOP_NEWTON) return loop_launch_kron_newton(op, a, env);
    return loop_launch_kron_anderson(op, a, env);
}

extern "C" {

int sdfs_solve_sa(sdfs_op *op, const double *d_w_init, double tol, int64_t max_iter, double *d_w_out,
                  int64_t *iters, double *final_err, double *d_err_hist, int64_t hist_stride, int64_t hist_cap) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_solve_sa: NULL op");
    sdfs_ctx *ctx = op->ctx;
    ARG_CHECK(ctx, d_w_init && d_w_out && max_iter >= 0);
    ARG_CHECK(ctx, !d_err_hist || (hist_stride >= 1 && hist_cap >= 1));
    if (!ctx->coop_supported) return sdfs_set_error(ctx, SDFS_ERR_UNSUPPORTED, "device lacks cooperative launch");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(op_ensure_work(op, 16));
    LoopEnv env;
    TRY(build_env(op, &env));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(LoopStatus), ctx->stream));
    if (env.nranks == 1) {
        // reference-sized grids: whole solve in one CTA with P resident in shared memory
        const int handled = small_sa_try(op, d_w_init, tol, max_iter, d_w_out, d_err_hist, hist_stride, hist_cap);
        if (handled < 0) return handled;
        if (handled == 1) {
            LoopStatus *hs0 = (LoopStatus *)ctx->h_status;
            TRY(finish_loop(ctx, hs0, 0));
            if (iters) *iters = hs0->iters;
            if (final_err) *final_err = hs0->final_err;
            return SDFS_OK;
        }
    }
    SAArgs a{};
    a.w_init = d_w_init;
    a.w[0] = op->work + 2 * op->ldv;
    a.w[1] = op->work + 3 * op->ldv;
    a.w_out = d_w_out;
    a.tol = tol;
    a.max_iter = max_iter;
    a.err_hist = d_err_hist;
    a.hist_stride = hist_stride > 0 ? hist_stride : 1;
    a.hist_cap = d_err_hist ? hist_cap : 0;
    if (op->storage == SDFS_STORAGE_DENSE) TRY(loop_launch_dense(op, LOOP_SA, &a, &env));
    else if (op->storage == SDFS_STORAGE_CONT) TRY(loop_launch_cont(op, LOOP_SA, &a, &env));
    else TRY(loop_launch_kron(op, LOOP_SA, &a, &env));
    ctx->launches++;
    LoopStatus *hs = (LoopStatus *)ctx->h_status;
    TRY(finish_loop(ctx, hs, 0));
    if (env.nranks > 1) {
        *comm_epoch(ctx) = hs->epoch_end;   // every rank passed the same barriers
        TRY(op_allgather(op, d_w_out));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (iters) *iters = hs->iters;
    if (final_err) *final_err = hs->final_err;
    return SDFS_OK;
}

int sdfs_solve_anderson(sdfs_op *op, const double *d_w_init, double tol, int64_t max_iter, int history_size,
                        int mixing_frequency, double beta, double ridge, double *d_w_out, int64_t *iters,
                        double *final_err) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_solve_anderson: NULL op");
    sdfs_ctx *ctx = op->ctx;
    ARG_CHECK(ctx, d_w_init && d_w_out && max_iter >= 0);
    ARG_CHECK(ctx, history_size >= 2 && history_size <= AND_MAX_HIST && mixing_frequency >= 1);
    if (!ctx->coop_supported) return sdfs_set_error(ctx, SDFS_ERR_UNSUPPORTED, "device lacks cooperative launch");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t N = op_N(op);
    TRY(op_ensure_work(op, 16 + 2 * AND_MAX_HIST));
    LoopEnv env;
    TRY(build_env(op, &env));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(LoopStatus), ctx->stream));
    AndersonArgs a{};
    const int64_t ld = op->ldv;
    a.w_init = d_w_init; a.w_out = d_w_out;
    a.x = op->work + 2 * ld; a.fx = op->work + 3 * ld;
    a.X = op->work + 16 * ld; a.R = op->work + (16 + AND_MAX_HIST) * ld;
    a.ldv = ld; a.tol = tol; a.max_iter = max_iter; a.m = history_size; a.mix = mixing_frequency;
    a.beta_mix = beta; a.ridge = ridge;
    if (op->storage == SDFS_STORAGE_DENSE) TRY(loop_launch_dense(op, LOOP_ANDERSON, &a, &env));
    else if (op->storage == SDFS_STORAGE_CONT) TRY(loop_launch_cont(op, LOOP_ANDERSON, &a, &env));
    else TRY(loop_launch_kron(op, LOOP_ANDERSON, &a, &env));
    ctx->launches++;
    LoopStatus *hs = (LoopStatus *)ctx->h_status;
    TRY(finish_loop(ctx, hs, 0));
    if (env.nranks > 1) {
        *comm_epoch(ctx) = hs->epoch_end;
        TRY(op_allgather(op, d_w_out));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (iters) *iters = hs->iters;
    if (final_err) *final_err = hs->final_err;
    return SDFS_OK;
}

int sdfs_solve_newton(sdfs_op *op, const double *d_w_init, double tol, int64_t max_iter, int krylov, double rtol,
                      double atol, int restart, int64_t krylov_maxiter, double *d_w_out, int64_t *outer_iters,
                      double *final_err, double *h_outer_err, int64_t *h_inner_iters, int64_t cap,
                      int64_t *total_matvecs) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_solve_newton: NULL op");
    sdfs_ctx *ctx = op->ctx;
    ARG_CHECK(ctx, d_w_init && d_w_out && max_iter >= 0);
    ARG_CHECK(ctx, krylov == SDFS_KRYLOV_BICGSTAB || krylov == SDFS_KRYLOV_GMRES);
    if (krylov == SDFS_KRYLOV_GMRES) ARG_CHECK(ctx, restart >= 1 && restart <= GMRES_MAX_RESTART);
    if (!ctx->coop_supported) return sdfs_set_error(ctx, SDFS_ERR_UNSUPPORTED, "device lacks cooperative launch");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t N = op_N(op);
    const int nvec = 16 + (krylov == SDFS_KRYLOV_GMRES ? restart + 1 : 0);
    TRY(op_ensure_work(op, nvec));
    LoopEnv env;
    TRY(build_env(op, &env));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(LoopStatus), ctx->stream));
    NewtonArgs a{};
    double *wk = op->work;
    const int64_t ld = op->ldv;
    a.w_init = d_w_init; a.w_out = d_w_out;
    a.w = wk + 2 * ld; a.g = wk + 3 * ld; a.c = wk + 4 * ld; a.d = wk + 5 * ld; a.x = wk + 6 * ld;
    a.r = wk + 7 * ld; a.rhat = wk + 8 * ld; a.p = wk + 9 * ld; a.q = wk + 10 * ld; a.s = wk + 11 * ld;
    a.t = wk + 12 * ld;
    a.V = wk + 16 * ld;
    a.ldv = ld;
    a.tol = tol; a.max_iter = max_iter; a.krylov = krylov; a.rtol = rtol; a.atol = atol;
    a.restart = restart;
    a.krylov_maxiter = krylov_maxiter > 0 ? krylov_maxiter : 10 * N;
    if (op->storage == SDFS_STORAGE_DENSE) TRY(loop_launch_dense(op, LOOP_NEWTON, &a, &env));
    else if (op->storage == SDFS_STORAGE_CONT) TRY(loop_launch_cont(op, LOOP_NEWTON, &a, &env));
    else TRY(loop_launch_kron(op, LOOP_NEWTON, &a, &env));
    ctx->launches++;
    LoopStatus *hs = (LoopStatus *)ctx->h_status;
    TRY(finish_loop(ctx, hs, 0));
    if (env.nranks > 1) {
        *comm_epoch(ctx) = hs->epoch_end;   // every rank passed the same barriers
        TRY(op_allgather(op, d_w_out));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (outer_iters) *outer_iters = hs->iters;
    if (final_err) *final_err = hs->final_err;
    if (total_matvecs) *total_matvecs = hs->matvecs;
    const int64_t nrec = hs->iters < HIST_CAP ? hs->iters : HIST_CAP;
    for (int64_t i = 0; i < cap && i < nrec; ++i) {
        if (h_outer_err) h_outer_err[i] = hs->outer_err[i];
        if (h_inner_iters) h_inner_iters[i] = hs->inner_iters[i];
    }
    return SDFS_OK;
}

}  // extern "C"
