// Factor-form (Kronecker) instantiation of one loop kernel per translation unit: these are the
// slowest units to compile (tensor-core contraction variants inlined at every apply site).
#pragma once
#include "loops.cuh"

template <int WHICH>
static int loop_launch_kron_t(sdfs_op *op, void *a, LoopEnv *env) {
    sdfs_ctx *ctx = op->ctx;
    const int64_t N = op->kv.N;
    if (!op->kron_tmp[0]) {
        CUDA_TRY(ctx, cudaMalloc(&op->kron_tmp[0], (size_t)N * sizeof(double)));
        CUDA_TRY(ctx, cudaMalloc(&op->kron_tmp[1], (size_t)N * sizeof(double)));
    }
    KronLoopOp lop{op->kvs, op->kron_tmp[0], op->kron_tmp[1]};     // this rank's (possibly slab-restricted) view
    const int64_t nloc = op->kvs.row_end - op->kvs.row_begin;
    // every rank must launch (the barriers span all ranks): a rank without rows still gets one CTA
    return loop_launch<KronLoopOp, WHICH>(ctx, lop, a, env, lop.dyn_smem(), 2, (nloc + KronLoopOp::kThreads - 1) / KronLoopOp::kThreads,
                                           op->kron_sharded /* every rank the same grid: the reduction slots are indexed by CTA */);
}
int loop_launch_kron_sa(sdfs_op *op, void *a, LoopEnv *env);
int loop_launch_kron_newton(sdfs_op *op, void *a, LoopEnv *env);
int loop_launch_kron_anderson(sdfs_op *op, void *a, LoopEnv *env);
