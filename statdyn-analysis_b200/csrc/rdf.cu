// rdf.cu -- pair-distance histogram of freud.density.RDF, the only heavy part of
// dynamics/_util.py:66-86 `calculate_wave_number` (Relaxations(..., wave_number=None),
// relaxations.py:61-62).  The reference asks for rmax = min(L)/2.2, i.e. nearly every pair of the box:
// the work is O(N^2) whatever the data structure, so this is a tiled all-pairs kernel.
//
//   for every ordered pair i != j:  r_ij = box.wrap(r_j - r_i)  (f32, freud operation order)
//       rsq = (x x + y y) + z z ;  rsq < rmax^2  ->  bin = floor(sqrt(rsq) / dr) ; bin < nbins -> ++count[bin]
//
// Counts are exact integers: bit-identical to oracle/freud_shim.py RDF.  The normalisation (shell
// volumes, density) and the structure-factor scan over 200 wave numbers are done on the host
// (sdanalysis_b200/dynamics/_util.py), 200 bins of work.
#include "sdb_internal.cuh"

namespace {

constexpr int RDF_THREADS = 256;
constexpr int RDF_TILE = 1024;      // j positions staged in shared memory per iteration
constexpr int RDF_MAX_BINS = 1024;

struct RdfArgs {
    SdbBox3 box;
    const float* pos;  // (n, 3)
    int n;
    float rmaxsq, dr_inv;
    int nbins;
    int j_per_block;   // j range handled by one blockIdx.y
    unsigned long long* counts;
};

// grid = (i tiles of 256, j chunks).  Every warp owns a private histogram in shared memory.
__global__ void __launch_bounds__(RDF_THREADS) rdf_hist_kernel(const RdfArgs a)
{
    extern __shared__ unsigned int hist[];  // [8 warps][nbins]
    __shared__ float3 tile[RDF_TILE];
    const int warp = threadIdx.x >> 5;
    unsigned int* mine = hist + warp * a.nbins;
    for (int b = threadIdx.x; b < (RDF_THREADS / 32) * a.nbins; b += RDF_THREADS) hist[b] = 0u;
    const int i = blockIdx.x * RDF_THREADS + threadIdx.x;
    float3 me = make_float3(0.f, 0.f, 0.f);
    if (i < a.n) me = make_float3(a.pos[3L * i], a.pos[3L * i + 1], a.pos[3L * i + 2]);
    const int j_begin = blockIdx.y * a.j_per_block;
    const int j_end = min(a.n, j_begin + a.j_per_block);
    for (int j0 = j_begin; j0 < j_end; j0 += RDF_TILE) {
        __syncthreads();
        const int count = min(RDF_TILE, j_end - j0);
        for (int t = threadIdx.x; t < count; t += RDF_THREADS) {
            const long j = j0 + t;
            tile[t] = make_float3(a.pos[3 * j], a.pos[3 * j + 1], a.pos[3 * j + 2]);
        }
        __syncthreads();
        if (i >= a.n) continue;
        for (int t = 0; t < count; ++t) {
            if (j0 + t == i) continue;
            const float3 o = tile[t];
            float x, y, z;
            sdb_wrap3(a.box, __fsub_rn(o.x, me.x), __fsub_rn(o.y, me.y), __fsub_rn(o.z, me.z), x, y, z);
            const float rsq = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
            if (rsq < a.rmaxsq) {
                const float binr = __fmul_rn(__fsqrt_rn(rsq), a.dr_inv);
                const unsigned int bin = (unsigned int)floorf(binr);
                if (bin < (unsigned int)a.nbins) atomicAdd(mine + bin, 1u);
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < a.nbins; b += RDF_THREADS) {
        unsigned long long total = 0;
        for (int w = 0; w < RDF_THREADS / 32; ++w) total += hist[w * a.nbins + b];
        if (total) atomicAdd(a.counts + b, total);
    }
}

}  // namespace

// Pair-distance histogram of freud.density.RDF(rmax, dr).compute(box, positions) (dynamics/_util.py:78-79).
// h_box: Lx, Ly, Lz, xy, xz, yz (freud Box); d_counts: uint64[nbins], overwritten.  nbins <= 1024.
extern "C" int sdb_rdf_histogram(const float* h_box, int32_t is2d, int32_t n, const float* d_pos, float rmax, float dr,
                                 int32_t nbins, uint64_t* d_counts, void* stream)
{
    if (!h_box || !d_pos || !d_counts || n <= 0 || nbins <= 0 || nbins > RDF_MAX_BINS || !(rmax > 0.f) || !(dr > 0.f)) {
        sdb_set_error("sdb_rdf_histogram: bad argument (n=%d nbins=%d)", n, nbins);
        return SDB_EINVAL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    RdfArgs a;
    a.box.lx = h_box[0];
    a.box.ly = h_box[1];
    a.box.lz = is2d ? 0.0f : h_box[2];
    a.box.xy = h_box[3];
    a.box.xz = is2d ? 0.0f : h_box[4];
    a.box.yz = is2d ? 0.0f : h_box[5];
    a.box.is2d = is2d;
    a.pos = d_pos;
    a.n = n;
    a.rmaxsq = rmax * rmax;   // f32 product, like freud
    a.dr_inv = 1.0f / dr;
    a.nbins = nbins;
    a.counts = reinterpret_cast<unsigned long long*>(d_counts);
    int rc = sdb_check_cuda(cudaMemsetAsync(d_counts, 0, sizeof(uint64_t) * (size_t)nbins, s), "sdb_rdf_histogram: clear");
    if (rc) return rc;
    const int i_blocks = (n + RDF_THREADS - 1) / RDF_THREADS;
    // enough j chunks to fill the GPU (148 SMs x 8 resident blocks) when there are few i tiles
    int j_chunks = (148 * 8 + i_blocks - 1) / i_blocks;
    const int max_chunks = (n + RDF_TILE - 1) / RDF_TILE;
    if (j_chunks > max_chunks) j_chunks = max_chunks;
    if (j_chunks < 1) j_chunks = 1;
    a.j_per_block = ((n + j_chunks - 1) / j_chunks + RDF_TILE - 1) / RDF_TILE * RDF_TILE;
    j_chunks = (n + a.j_per_block - 1) / a.j_per_block;
    const size_t smem = sizeof(unsigned int) * (RDF_THREADS / 32) * (size_t)nbins;
    rdf_hist_kernel<<<dim3(i_blocks, j_chunks), RDF_THREADS, smem, s>>>(a);
    SDB_CHECK_LAUNCH("rdf_hist_kernel");
    return SDB_OK;
}
