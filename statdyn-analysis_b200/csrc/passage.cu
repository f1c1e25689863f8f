// passage.cu -- stand-alone first/last-passage trackers on caller-provided distances.
//
// API parity for MolecularRelaxation.add (reference dynamics/relaxations.py:209-218) and
// LastMolecularRelaxation.add (:238-289) when they are driven directly with a distance array, as the
// reference tests do (test/dynamics_test.py:191-265).  The fused hot path (dyn_step.cu) keeps the
// same state machines in one flag byte per molecule; these kernels keep the reference's own layout
// (int64 status word, uint8 state).  Branch-free: every update is a select.
#include "sdb_internal.cuh"

namespace {

template <typename T>
__global__ void first_passage_kernel(const T* __restrict__ dist, int n, T threshold, long long timediff,
                                     long long* __restrict__ status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long st = status[i];
    // status[(dist > threshold) & (status > timediff)] = timediff ; NaN compares false
    const bool hit = (dist[i] > threshold) & (st > timediff);
    status[i] = hit ? timediff : st;
}

template <typename T>
__global__ void last_passage_kernel(const T* __restrict__ dist, int n, T threshold, T irreversibility,
                                    long long timediff, long long* __restrict__ status, uint8_t* __restrict__ state)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T d = dist[i];
    unsigned s = state[i];
    long long st = status[i];
    // the reference's three masked passes, in order (relaxations.py:268-289)
    const bool passed = (d > threshold) & (s == 0u);
    s = passed ? 1u : s;
    st = passed ? timediff : st;
    const bool below = (d < threshold) & (s == 1u);
    s = below ? 0u : s;
    const bool irreversible = (d > irreversibility) & (s == 1u);
    s = irreversible ? 2u : s;
    state[i] = (uint8_t)s;
    status[i] = st;
}

}  // namespace

extern "C" int sdb_first_passage(const void* d_dist, int32_t dist_is_f64, int32_t n, double threshold, int64_t timediff,
                                 int64_t* d_status, void* stream)
{
    if (!d_dist || !d_status || n <= 0) {
        sdb_set_error("sdb_first_passage: bad argument");
        return SDB_EINVAL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = (n + 255) / 256;
    if (dist_is_f64)
        first_passage_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(d_dist), n, threshold,
                                                          (long long)timediff, reinterpret_cast<long long*>(d_status));
    else
        first_passage_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(d_dist), n, (float)threshold,
                                                         (long long)timediff, reinterpret_cast<long long*>(d_status));
    SDB_CHECK_LAUNCH("first_passage_kernel");
    return SDB_OK;
}

extern "C" int sdb_last_passage(const void* d_dist, int32_t dist_is_f64, int32_t n, double threshold,
                                double irreversibility, int64_t timediff, int64_t* d_status, uint8_t* d_state,
                                void* stream)
{
    if (!d_dist || !d_status || !d_state || n <= 0) {
        sdb_set_error("sdb_last_passage: bad argument");
        return SDB_EINVAL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = (n + 255) / 256;
    if (dist_is_f64)
        last_passage_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(d_dist), n, threshold,
                                                         irreversibility, (long long)timediff,
                                                         reinterpret_cast<long long*>(d_status), d_state);
    else
        last_passage_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(d_dist), n, (float)threshold,
                                                        (float)irreversibility, (long long)timediff,
                                                        reinterpret_cast<long long*>(d_status), d_state);
    SDB_CHECK_LAUNCH("last_passage_kernel");
    return SDB_OK;
}
