// api.cu -- library-wide plumbing of libsdb200.so: error reporting, launch counter, the atan table,
// host-side test hooks.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "sdb_internal.cuh"

namespace {
thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};

SdbAtanTable g_host_table;
std::once_flag g_host_table_once;
__device__ SdbAtanTable g_dev_table;
std::mutex g_dev_mutex;
bool g_dev_ready[64] = {false};

void build_host_table()
{
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < 1024; ++i) {
        const double t = two_pi * (double)i / 1024.0;
        g_host_table.cs[i][0] = cos(t);
        g_host_table.cs[i][1] = sin(t);
    }
    // exact values on the axes
    g_host_table.cs[0][0] = 1.0;    g_host_table.cs[0][1] = 0.0;
    g_host_table.cs[256][0] = 0.0;  g_host_table.cs[256][1] = 1.0;
    g_host_table.cs[512][0] = -1.0; g_host_table.cs[512][1] = 0.0;
    g_host_table.cs[768][0] = 0.0;  g_host_table.cs[768][1] = -1.0;
}
}  // namespace

void sdb_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sdb_check_cuda(cudaError_t err, const char* what)
{
    if (err == cudaSuccess) return SDB_OK;
    sdb_set_error("%s: %s", what, cudaGetErrorString(err));
    return SDB_ECUDA;
}

void sdb_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

const double (*sdb_host_atan_table())[2]
{
    std::call_once(g_host_table_once, build_host_table);
    return g_host_table.cs;
}

const double (*sdb_device_atan_table())[2]
{
    int dev = 0;
    if (sdb_check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return nullptr;
    void* sym = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        if (dev < 64 && !g_dev_ready[dev]) {
            sdb_host_atan_table();
            if (sdb_check_cuda(cudaMemcpyToSymbol(g_dev_table, &g_host_table, sizeof(SdbAtanTable)), "atan table upload"))
                return nullptr;
            g_dev_ready[dev] = true;
        }
    }
    if (sdb_check_cuda(cudaGetSymbolAddress(&sym, g_dev_table), "cudaGetSymbolAddress")) return nullptr;
    return reinterpret_cast<const double (*)[2]>(sym);
}

extern "C" const char* sdb_last_error(void) { return g_error; }
extern "C" const char* sdb_version(void) { return "sdb200 0.1.0 (sm_100a)"; }
extern "C" uint64_t sdb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" double sdb_test_rot_increment(float w, float z, float pw, float pz, int32_t* bad)
{
    int b = 0;
    const double r = sdb_rot_increment(w, z, pw, pz, sdb_host_atan_table(), b);
    if (bad) *bad = b;
    return r;
}

extern "C" void sdb_test_wrap2d(float lx, float ly, float xy, float vx, float vy, float* ox, float* oy)
{
    sdb_wrap2d(lx, ly, xy, vx, vy, *ox, *oy);
}

// Page-lock a host range that the caller keeps mapped (a memory-mapped trajectory file): frames inside it are then
// uploaded straight from the page cache.  Read-only registration first (the mapping is usually PROT_READ), then the
// default flavour; a failure is not an error of the library (the caller falls back to staged copies) and leaves no
// pending CUDA error behind.
extern "C" int sdb_host_register(void* h_ptr, size_t bytes)
{
    if (!h_ptr || bytes == 0) return SDB_EINVAL;
    cudaError_t err = cudaHostRegister(h_ptr, bytes, cudaHostRegisterReadOnly);
    if (err != cudaSuccess) {
        cudaGetLastError();
        err = cudaHostRegister(h_ptr, bytes, cudaHostRegisterDefault);
    }
    if (err != cudaSuccess) {
        cudaGetLastError();
        sdb_set_error("sdb_host_register: %s", cudaGetErrorString(err));
        return SDB_ECUDA;
    }
    return SDB_OK;
}

extern "C" int sdb_host_unregister(void* h_ptr)
{
    if (!h_ptr) return SDB_EINVAL;
    const cudaError_t err = cudaHostUnregister(h_ptr);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return SDB_ECUDA;
    }
    return SDB_OK;
}

// Asynchronous host-to-device copy on `stream` (page-locked source: truly asynchronous; pageable source: the
// runtime stages it, blocking the host but not synchronising the stream's earlier work away from the caller).
extern "C" int sdb_upload_async(void* d_dst, const void* h_src, size_t bytes, void* stream)
{
    if (!d_dst || !h_src) {
        sdb_set_error("sdb_upload_async: null argument");
        return SDB_EINVAL;
    }
    if (bytes == 0) return SDB_OK;
    return sdb_check_cuda(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)),
                          "sdb_upload_async");
}

// ---- small copies on the SMs ------------------------------------------------------------------------------------
// The per-frame descriptor upload (a few KB, host -> device) and the per-frame row write into another GPU's memory
// (a few KB, peer) are normally cudaMemcpyAsync calls, i.e. they are served by the copy engines in submission order.
// When a frame slice of tens of MB is being uploaded at the same time (multi-GPU end-to-end runs: every rank
// uploads its slice on a side stream), those tiny copies queue behind it and the compute stream starts its next
// frame up to a slice-upload late (measured on 2 GPUs: e2e 11 % below `value`; with the copies on the SMs 3 %).
// The native engine therefore moves them with one block of 16-byte word copies (the source may be mapped pinned
// host memory, the destination peer memory); SDB_SMALL_COPIES=engine goes back to cudaMemcpyAsync.
namespace {
__global__ void __launch_bounds__(256) small_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, int n_vec)
{
    for (int i = threadIdx.x; i < n_vec; i += blockDim.x) dst[i] = src[i];
}
}  // namespace

bool sdb_small_copies_on_sm()
{
    static const bool on = [] {
        const char* e = getenv("SDB_SMALL_COPIES");
        return !(e && e[0] == 'e');  // "engine": back to cudaMemcpyAsync
    }();
    return on;
}

// bytes is rounded up to 16: both buffers must have that much room
int sdb_small_copy(void* dst, const void* src, size_t bytes, void* stream)
{
    if (bytes == 0) return SDB_OK;
    const int n_vec = (int)((bytes + 15) / 16);
    small_copy_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src), n_vec);
    SDB_CHECK_LAUNCH("small_copy_kernel");
    return SDB_OK;
}
