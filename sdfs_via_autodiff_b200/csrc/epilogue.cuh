// Epilogues of one operator application and the fused result exchange of sharded operators (shared by the
// dense pass, ops.cu, and the factor-form apply, kron_apply.cu).
#pragma once
#include "common.cuh"
#include "arena.cuh"

// Epilogue modes of the dense / factor-form pass
//  0: out0 = 1 + beta (a_row s0)^(1/theta)                                   (T)
//  1: out0 = beta (a_row s0)^((1-theta)/theta) a_row s1                      (J_T(w) v)
//  2: out0 = beta^theta e_sdf (w-1)^(1-theta) s1 ; out1 = beta^theta a_row s0/(w-1)^theta - 1 (SDF)
//  3: out0 = s0                                                              (P x)
struct EpiArgs {
    int mode;
    const double *a_row, *e_sdf, *w;
    double beta, theta;
    double *out0, *out1;
    double inv_theta = 0.0;     // 1 / theta, filled by the launchers
};

template <bool FAST = false>
__device__ __forceinline__ void apply_epilogue(const EpiArgs &e, int64_t n, double s0, double s1) {
    auto pw = [](double x, double ex) { return FAST ? pow_pos(x, ex) : pow(x, ex); };
    if (e.mode == 0) {
        e.out0[n] = 1.0 + e.beta * pw(e.a_row[n] * s0, e.inv_theta);
    } else if (e.mode == 1) {
        const double ar = e.a_row[n];
        e.out0[n] = e.beta * pw(ar * s0, e.inv_theta - 1.0) * ar * s1;
    } else if (e.mode == 2) {
        const double bt = pow(e.beta, e.theta);
        const double wm1 = e.w[n] - 1.0;
        if (e.out0) e.out0[n] = bt * e.e_sdf[n] * pw(wm1, 1.0 - e.theta) * s1;
        if (e.out1) e.out1[n] = bt * (e.a_row[n] * s0) / pw(wm1, e.theta) - 1.0;
    } else {
        e.out0[n] = s0;
    }
}

// single-output epilogues (T, JVP, plain P x) as a value, for the fused exchange below
template <bool FAST = false>
__device__ __forceinline__ double epilogue_value(const EpiArgs &e, int64_t n, double s0, double s1) {
    auto pw = [](double x, double ex) { return FAST ? pow_pos(x, ex) : pow(x, ex); };
    if (e.mode == 0) return 1.0 + e.beta * pw(e.a_row[n] * s0, e.inv_theta);
    if (e.mode == 1) {
        const double ar = e.a_row[n];
        return e.beta * pw(ar * s0, e.inv_theta - 1.0) * ar * s1;
    }
    return s0;
}

// Fused exchange of a row-sharded single application (one process per GPU): the epilogue stores its
// rows straight into the result buffer of EVERY rank (NVLink peer stores into the CUDA-IPC arenas),
// and the last CTA of the kernel to finish trades an epoch flag with the peers, so when the kernel
// ends the full vector is in this rank's arena - no collective launch follows the row pass.
struct PeerArgs {
    int nranks, rank;                          // nranks <= 1: plain local stores through EpiArgs
    double *out[SDFS_MAX_RANKS];               // result buffer y[epoch & 1] in rank r's arena
    unsigned long long *sig[SDFS_MAX_RANKS];   // rank r's flag word for this rank
    unsigned long long *mine;                  // this rank's flag words (written by the peers)
    unsigned long long epoch;
    unsigned int *counter;                     // CTAs of this launch that have finished (self-resetting)
    long long *h_abort;                        // pinned host word: set when a peer never arrives
};

__device__ __forceinline__ void peer_exchange_finish(const PeerArgs &pa) {
    __syncthreads();                           // every store of this CTA issued
    if (threadIdx.x == 0) {
        __threadfence_system();                // ... and ordered before the arrival count
        const unsigned int done = atomicAdd(pa.counter, 1u);
        if (done == gridDim.x - 1) {           // last CTA of this rank
            *pa.counter = 0;
            __threadfence_system();
            for (int r = 0; r < pa.nranks; ++r)
                if (r != pa.rank) st_release_sys(pa.sig[r], pa.epoch);
            const long long t0 = clock64();
            for (int r = 0; r < pa.nranks; ++r) {
                if (r == pa.rank) continue;
                while (ld_acquire_sys(pa.mine + r) < pa.epoch) {
                    if (clock64() - t0 > SDFS_PEER_TIMEOUT_CLOCKS) {
                        *(volatile long long *)pa.h_abort = 1;
                        __threadfence_system();
                        return;
                    }
                }
            }
        }
    }
}

