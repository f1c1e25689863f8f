// exchange.cu -- frame exchange of the origin-sharded multi-GPU path (no reference counterpart: the
// reference is a single process).
//
// Every rank owns a disjoint set of time-origins; the only shared input is the current frame
// (positions + quaternions, 28 B per molecule).  The reader rank pushes it into the same slot of every
// rank's frame ring with ONE pass of multimem.st stores through an NVSwitch multicast address: the
// switch replicates the stores, so the reader's NVLink egress is 28 B per molecule whatever the
// number of ranks (a ring/tree broadcast forwards every byte R-1 times through SMs instead).
#include "sdb_internal.cuh"

namespace {

// 16-byte loads from local HBM, 16-byte multicast stores.  A small grid is enough to fill one
// NVLink port set (~770 GB/s); it runs on a side stream next to the compute kernels.
__global__ void __launch_bounds__(256) multicast_copy_kernel(const float4* __restrict__ src, float* mc_dst, size_t n_vec)
{
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += step) {
        const float4 v = __ldcs(src + i);
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + 4 * i), "f"(v.x), "f"(v.y),
                     "f"(v.z), "f"(v.w)
                     : "memory");
    }
}

}  // namespace

extern "C" int sdb_multicast_copy(const void* d_src, void* d_multicast_dst, size_t bytes, int32_t blocks, void* stream)
{
    if (!d_src || !d_multicast_dst || (bytes & 15) || ((uintptr_t)d_src & 15) || ((uintptr_t)d_multicast_dst & 15)) {
        sdb_set_error("sdb_multicast_copy: pointers and size must be 16-byte aligned");
        return SDB_EINVAL;
    }
    if (bytes == 0) return SDB_OK;
    if (blocks <= 0) blocks = 64;
    multicast_copy_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float4*>(d_src), static_cast<float*>(d_multicast_dst), bytes / 16);
    SDB_CHECK_LAUNCH("multicast_copy_kernel");
    return SDB_OK;
}
