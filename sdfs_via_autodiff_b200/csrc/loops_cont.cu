// Loop kernels (SA, Newton, Anderson) instantiated for the continuous-state operator.
#include "loops.cuh"

int loop_launch_cont(sdfs_op *op, int which, void *a, LoopEnv *env) {
    ContLoopOp lop{op->cv};
    const int64_t groups = (op->cv.N + 8) / 9;
    if (which == LOOP_SA) return loop_launch<ContLoopOp, LOOP_SA>(op->ctx, lop, a, env, 0, 2, groups, false);
    if (which == LOOP_NEWTON) return loop_launch<ContLoopOp, LOOP_NEWTON>(op->ctx, lop, a, env, 0, 2, groups, false);
    return loop_launch<ContLoopOp, LOOP_ANDERSON>(op->ctx, lop, a, env, 0, 2, groups, false);
}
