// Loop kernels (SA, Newton, Anderson) instantiated for the dense row-streamed operator.
#include "loops.cuh"

int loop_launch_dense(sdfs_op *op, int which, void *a, LoopEnv *env) {
    DenseLoopOp lop{op->dv};
    const size_t smem = lop.dyn_smem();
    const int64_t groups = dense_groups(op->dv);
    const bool full = env->nranks > 1;
    if (which == LOOP_SA) return loop_launch<DenseLoopOp, LOOP_SA>(op->ctx, lop, a, env, smem, 1, groups, full);
    if (which == LOOP_NEWTON) return loop_launch<DenseLoopOp, LOOP_NEWTON>(op->ctx, lop, a, env, smem, 1, groups, full);
    return loop_launch<DenseLoopOp, LOOP_ANDERSON>(op->ctx, lop, a, env, smem, 1, groups, full);
}
