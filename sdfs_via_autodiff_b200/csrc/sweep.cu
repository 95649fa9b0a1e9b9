// Batched (gamma, psi, beta) parameter sweep -- placeholder until the fp64
// tensor-core GEMM lands (see DESIGN.md, section "Sweep").
#include "common.cuh"

extern "C" {

int sdfs_sweep_solve_sa(sdfs_op *op, const double *, int64_t, double, double, int64_t, double *, int64_t *, double *) {
    return sdfs_set_error(op ? op->ctx : nullptr, SDFS_ERR_UNSUPPORTED, "sweep: not built in this revision");
}

int sdfs_sweep_apply_T(sdfs_op *op, const double *, int64_t, const double *, double *) {
    return sdfs_set_error(op ? op->ctx : nullptr, SDFS_ERR_UNSUPPORTED, "sweep: not built in this revision");
}

}  // extern "C"
