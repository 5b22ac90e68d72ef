// f64 estimators (frequency offset, symbol timing): fused filter + product + reduction kernels.
#pragma once
#include "common.cuh"

namespace cb {

unsigned estimator_max_partials();  // entries the `partial` scratch must hold
// out[0] = sum_{i < n-1} x[i+1] conj(x[i])
int launch_freq_sum(const double2 *x, size_t n, double2 *partial, double2 *out, cudaStream_t s);
// out[0] = sum_i qout[i] dout[i] of TimingEstimator::push; taps = the ntaps = 2 nd + 1 values of q(t), sps = N
size_t timing_smem_bytes(unsigned ntaps);
int launch_timing_sum(const double2 *x, size_t n, const double *taps, unsigned ntaps, unsigned nd, unsigned sps,
                      double2 *partial, double2 *out, cudaStream_t s);

}  // namespace cb
