// sdb_internal.cuh -- shared declarations of the library's translation units.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>

#include "../../include/sdb200.h"
#include "sdb_math.cuh"

// ---- tiling of the fused dynamics kernel ------------------------------------------------------
// A warp owns 256 consecutive molecules: two groups of 128, each lane 4 consecutive molecules per
// group, so every per-origin plane (dx, dy, dtheta: f32; flags: u8) is read and written with fully
// coalesced 16-byte (4-byte for flags) accesses.  Warps never synchronise with each other.
constexpr int SDB_LANE_MOLS = 4;
constexpr int SDB_GROUPS = 2;
constexpr int SDB_WARP_MOLS = 32 * SDB_LANE_MOLS * SDB_GROUPS;  // 256
constexpr int SDB_BLOCK_WARPS = 4;
constexpr int SDB_BLOCK_THREADS = 32 * SDB_BLOCK_WARPS;
constexpr int SDB_NFSUM = 10;  // fp64 sums reduced through shared memory (SDB_S_DISP..SDB_S_D2R2)

void sdb_set_error(const char* fmt, ...);
int sdb_check_cuda(cudaError_t err, const char* what);
void sdb_count_launch(int n = 1);
const double (*sdb_device_atan_table())[2];  // device pointer to the 1024-entry cos/sin table
const double (*sdb_host_atan_table())[2];

#define SDB_CHECK_LAUNCH(what)                                       \
    do {                                                             \
        sdb_count_launch();                                          \
        int _rc = sdb_check_cuda(cudaGetLastError(), what);          \
        if (_rc) return _rc;                                         \
    } while (0)

__device__ __forceinline__ float4 sdb_ldg4(const float* p)
{
    return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming (evict-first) 16-byte accesses for state that is touched once per frame
__device__ __forceinline__ float4 sdb_ld_stream4(const float* p)
{
    return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void sdb_st_stream4(float* p, float4 v)
{
    __stcs(reinterpret_cast<float4*>(p), v);
}
