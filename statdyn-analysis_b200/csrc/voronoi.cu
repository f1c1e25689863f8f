// voronoi.cu -- number of Voronoi neighbours of every molecule in a periodic 2D box: replaces
// freud.voronoi.Voronoi(box, buff).computeNeighbors(position).nlist.neighbor_counts as used by the
// reference's order.compute_voronoi_neighs (order.py:312-315).  freud builds periodic buffer images and
// hands them to qhull; here every molecule's cell is built directly:
//
//   the cell of i starts as a square around i and is clipped by the bisector of (i, j) for the
//   neighbours j in order of distance (the k <= 16 nearest within rmax from sdb_neighbours); once the
//   next neighbour is farther than twice the farthest cell vertex no other molecule can cut the cell
//   ("security radius") and the number of edges is the answer.  A molecule whose candidate list ends
//   before that point is redone against all molecules by a second kernel (rare: voids).
//
// Differences r_j - r_i of the f32 inputs are exact in f64, the minimum image (orthorhombic box) and
// the clipping are f64: for non-degenerate inputs the counts are those of the exact Voronoi diagram.
#include "sdb_internal.cuh"

namespace {

constexpr int VOR_MAXV = 40;         // first kernel: neighbours arrive in order of distance, cells stay small
constexpr int VOR_MAXV_BRUTE = 128;  // second kernel: arbitrary order inside a shell
constexpr uint32_t VOR_MISSING = 0xFFFFFFFFu;

struct VorArgs {
    const float* pos;  // (n, 3)
    int n;
    double lx, ly;
    const uint32_t* nlist;  // (n, k) sorted by distance, VOR_MISSING padded
    int k;
    double rmax;
    uint32_t* counts;       // (n)
    uint32_t* unresolved;   // [0] = count, [1..] = molecule indices
};

template <int CAP>
struct PolyT {
    double x[CAP], y[CAP];
    int tag[CAP];  // generator of the edge leaving vertex v (-1: the initial square)
    int nv;
};
using Poly = PolyT<VOR_MAXV>;

template <int CAP>
__device__ __forceinline__ void poly_square(PolyT<CAP>& p, double b)
{
    p.nv = 4;
    p.x[0] = -b; p.y[0] = -b;
    p.x[1] = b;  p.y[1] = -b;
    p.x[2] = b;  p.y[2] = b;
    p.x[3] = -b; p.y[3] = b;
    for (int v = 0; v < 4; ++v) p.tag[v] = -1;
}

// Keep the part of the convex polygon with x . d <= |d|^2 / 2 (closer to the origin than to d).
// Returns false when the vertex buffer would overflow.
template <int CAP>
__device__ bool poly_clip(const PolyT<CAP>& in, PolyT<CAP>& out, double dx, double dy, int newtag)
{
    const double c = 0.5 * (dx * dx + dy * dy);
    int m = 0;
    for (int v = 0; v < in.nv; ++v) {
        const int w = v + 1 == in.nv ? 0 : v + 1;
        const double sa = in.x[v] * dx + in.y[v] * dy - c, sb = in.x[w] * dx + in.y[w] * dy - c;
        const bool ina = sa <= 0.0, inb = sb <= 0.0;
        if (ina) {
            if (m >= CAP) return false;
            out.x[m] = in.x[v]; out.y[m] = in.y[v]; out.tag[m] = in.tag[v];
            ++m;
        }
        if (ina != inb) {
            if (m >= CAP) return false;
            const double t = sa / (sa - sb);
            out.x[m] = in.x[v] + t * (in.x[w] - in.x[v]);
            out.y[m] = in.y[v] + t * (in.y[w] - in.y[v]);
            out.tag[m] = ina ? newtag : in.tag[v];  // leaving the half plane: the new edge starts here
            ++m;
        }
    }
    out.nv = m;
    return true;
}

template <int CAP>
__device__ __forceinline__ double poly_radius2(const PolyT<CAP>& p)
{
    double r2 = 0.0;
    for (int v = 0; v < p.nv; ++v) r2 = fmax(r2, p.x[v] * p.x[v] + p.y[v] * p.y[v]);
    return r2;
}

template <int CAP>
__device__ __forceinline__ bool poly_closed(const PolyT<CAP>& p)
{
    for (int v = 0; v < p.nv; ++v)
        if (p.tag[v] < 0) return false;
    return p.nv >= 3;
}

__device__ __forceinline__ void min_image(const VorArgs& a, int i, int j, double& dx, double& dy)
{
    dx = (double)a.pos[3L * j] - (double)a.pos[3L * i];
    dy = (double)a.pos[3L * j + 1] - (double)a.pos[3L * i + 1];
    dx -= a.lx * rint(dx / a.lx);
    dy -= a.ly * rint(dy / a.ly);
}

__global__ void __launch_bounds__(128) voronoi_kernel(const VorArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    Poly buf[2];
    int cur = 0;
    poly_square(buf[0], a.rmax);
    double r2 = 2.0 * a.rmax * a.rmax;  // farthest vertex, squared
    bool ok = true, secured = false;
    double last_d2 = 0.0;
    int used = 0;
    for (int s = 0; s < a.k; ++s) {
        const uint32_t j = a.nlist[(long)i * a.k + s];
        if (j == VOR_MISSING) break;
        double dx, dy;
        min_image(a, i, (int)j, dx, dy);
        const double d2 = dx * dx + dy * dy;
        last_d2 = d2;
        ++used;
        if (d2 > 4.0 * r2) {  // no molecule at this distance or beyond can cut the cell
            secured = true;
            break;
        }
        if (!poly_clip(buf[cur], buf[cur ^ 1], dx, dy, s)) {
            ok = false;
            break;
        }
        cur ^= 1;
        r2 = poly_radius2(buf[cur]);
    }
    if (ok && !secured) {
        // the list ended: everything within `reach` has been seen (rmax if the list is not full)
        const double reach2 = used < a.k ? a.rmax * a.rmax : last_d2;
        secured = reach2 >= 4.0 * r2;
    }
    if (ok && secured && poly_closed(buf[cur])) {
        a.counts[i] = (uint32_t)buf[cur].nv;
    } else {
        a.counts[i] = VOR_MISSING;
        const uint32_t slot = atomicAdd(a.unresolved, 1u);
        a.unresolved[1 + slot] = (uint32_t)i;
    }
}

// Unresolved molecules against every other molecule, in shells of doubling radius: within a shell the order
// is arbitrary (clipping commutes; a molecule farther than twice the current farthest vertex cannot cut
// the current polygon, which contains the cell), but going outwards shell by shell keeps the intermediate
// polygons small -- clipping by a far-away row of molecules first would give a cell with one edge per
// molecule of the row.  After a shell of outer radius R everything closer than R has been seen, so the cell
// is final once its farthest vertex is within R / 2.
// One WARP per unresolved molecule (round 1: one thread, which made a 32 k-molecule frame with a void strip three
// times slower than a 1 M-molecule one without): the lanes test 32 candidates at a time, the few that pass are
// clipped in candidate order by lane 0 on a polygon in shared memory -- the same sequence of clips as a serial scan.
constexpr int VOR_BRUTE_WARPS = 4;

__global__ void __launch_bounds__(32 * VOR_BRUTE_WARPS) voronoi_brute_kernel(const VorArgs a)
{
    __shared__ PolyT<VOR_MAXV_BRUTE> polys[VOR_BRUTE_WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_unresolved = a.unresolved[0];
    const double b = 0.5 * fmin(a.lx, a.ly);  // beyond this a cell would meet its own periodic images
    for (uint32_t u = blockIdx.x * VOR_BRUTE_WARPS + warp; u < n_unresolved; u += gridDim.x * VOR_BRUTE_WARPS) {
        const int i = (int)a.unresolved[1 + u];
        PolyT<VOR_MAXV_BRUTE>* buf = polys[warp];
        int cur = 0;
        if (lane == 0) poly_square(buf[0], b);
        __syncwarp();
        double r2 = 2.0 * b * b;
        bool ok = true;
        double rin2 = 0.0, rout = a.rmax;
        while (ok) {
            const double rout2 = rout * rout;
            for (int base = 0; base < a.n && ok; base += 32) {
                const int j = base + lane;
                double dx = 0.0, dy = 0.0, d2 = INFINITY;
                if (j < a.n && j != i) {
                    min_image(a, i, j, dx, dy);
                    d2 = dx * dx + dy * dy;
                }
                unsigned mask = __ballot_sync(0xffffffffu, d2 >= rin2 && d2 < rout2 && d2 < 4.0 * r2);
                while (mask && ok) {  // (warp-uniform)
                    const int l = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const double cx = __shfl_sync(0xffffffffu, dx, l), cy = __shfl_sync(0xffffffffu, dy, l);
                    if (cx * cx + cy * cy >= 4.0 * r2) continue;  // the cell has shrunk since the test above
                    int clipped = 1;
                    double nr2 = r2;
                    if (lane == 0) {
                        clipped = poly_clip(buf[cur], buf[cur ^ 1], cx, cy, base + l) ? 1 : 0;
                        if (clipped) nr2 = poly_radius2(buf[cur ^ 1]);
                    }
                    clipped = __shfl_sync(0xffffffffu, clipped, 0);
                    nr2 = __shfl_sync(0xffffffffu, nr2, 0);
                    ok = clipped != 0;
                    cur ^= 1;
                    r2 = nr2;
                }
            }
            if (4.0 * r2 <= rout2 || rout2 >= 2.0 * b * b) break;
            rin2 = rout2;
            rout *= 2.0;
        }
        // a cell that still touches the starting square is bounded by the molecule's own periodic images:
        // not representable here (freud's buffer would not cover it either); reported as VOR_MISSING
        if (lane == 0) a.counts[i] = ok && poly_closed(buf[cur]) ? (uint32_t)buf[cur].nv : VOR_MISSING;
        __syncwarp();
    }
}

}  // namespace

// Voronoi neighbour counts (order.py:312-315).  h_box: Lx, Ly, Lz, xy, xz, yz of a 2D orthorhombic box;
// d_nlist: (n, k) nearest neighbours within rmax sorted by distance (sdb_neighbours), k <= 16;
// d_counts: uint32[n] out (0xFFFFFFFF: the cell could not be closed); d_scratch: uint32[n + 1].
extern "C" int sdb_voronoi_counts(const float* h_box, int32_t n, const float* d_pos, const uint32_t* d_nlist, int32_t k,
                                  float rmax, uint32_t* d_counts, uint32_t* d_scratch, void* stream)
{
    if (!h_box || !d_pos || !d_nlist || !d_counts || !d_scratch || n <= 0 || k <= 0 || k > SDB_MAX_K || !(rmax > 0.f)) {
        sdb_set_error("sdb_voronoi_counts: bad argument (n=%d k=%d)", n, k);
        return SDB_EINVAL;
    }
    if (h_box[3] != 0.0f) {
        sdb_set_error("sdb_voronoi_counts: tilted boxes are not supported");
        return SDB_EINVAL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VorArgs a;
    a.pos = d_pos;
    a.n = n;
    a.lx = (double)h_box[0];
    a.ly = (double)h_box[1];
    a.nlist = d_nlist;
    a.k = k;
    a.rmax = (double)rmax;
    a.counts = d_counts;
    a.unresolved = d_scratch;
    int rc = sdb_check_cuda(cudaMemsetAsync(d_scratch, 0, sizeof(uint32_t), s), "sdb_voronoi_counts: clear");
    if (rc) return rc;
    voronoi_kernel<<<(n + 127) / 128, 128, 0, s>>>(a);
    SDB_CHECK_LAUNCH("voronoi_kernel");
    // the second kernel sizes itself from the device-side count: a fixed grid of warps strides over the list
    voronoi_brute_kernel<<<148 * 4, 32 * VOR_BRUTE_WARPS, 0, s>>>(a);
    SDB_CHECK_LAUNCH("voronoi_brute_kernel");
    return SDB_OK;
}
