#include "loops_kron.cuh"

int loop_launch_kron_newton(sdfs_op *op, void *a, LoopEnv *env) { return loop_launch_kron_t<LOOP_NEWTON>(op, a, env); }
