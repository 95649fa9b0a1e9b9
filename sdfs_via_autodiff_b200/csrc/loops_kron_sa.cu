#include "loops_kron.cuh"

int loop_launch_kron_sa(sdfs_op *op, void *a, LoopEnv *env) { return loop_launch_kron_t<LOOP_SA>(op, a, env); }
