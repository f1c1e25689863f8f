/* sdb200.h -- C-ABI of libsdb200.so: B200 (sm_100a) kernels for the sdanalysis per-frame dynamics
 * hot path.  Plain C, no torch types.  Every pointer named d_* is a DEVICE pointer owned by the
 * caller (the Python host keeps them in torch tensors); h_* are HOST pointers.  All calls enqueue
 * work on the given stream (a cudaStream_t passed as void*) and return immediately unless stated
 * otherwise.  Return value: 0 on success, negative SDB_E* code on failure; sdb_last_error() gives
 * the message of the last failure on the calling thread.
 *
 * The reference (malramsay64/statdyn-analysis, sdanalysis v0.11.6) is pure Python: it has no FFI
 * for this path.  Each entry point below therefore names the reference *Python* interface whose
 * arithmetic it replaces (file:line relative to the reference checkout); INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 */
#ifndef SDB200_H
#define SDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_OK 0
#define SDB_EINVAL (-1) /* bad argument */
#define SDB_ECUDA (-2)  /* CUDA runtime error */

#define SDB_MAX_TRACKERS 6 /* relaxations.py:67-81 -- the six default trackers */
#define SDB_NO_STATUS 0xFFFFFFFFu /* relaxations.py:190,199 -- "not relaxed" = 2**32 - 1 */

/* Number of fp64 sums produced per (frame, origin) by sdb_dyn_step, see enum below. */
#define SDB_NSUM 16

/* Index of each per-(frame, origin) sum in the output rows of sdb_dyn_step / sdb_dyn_finalize.
 * The 15 quantities of Dynamics.compute_all (dynamics.py:49-65,328-394) are closed-form functions
 * of these sums and of N (done on the host, dynamics.py:183-303). */
enum {
    SDB_S_DISP = 0,    /* sum |dr|                    -> mean_displacement   dynamics.py:357-359 */
    SDB_S_DISP2 = 1,   /* sum |dr|^2                  -> msd                 dynamics.py:187-189 */
    SDB_S_DISP4 = 2,   /* sum |dr|^4                  -> mfd, alpha          dynamics.py:191-211 */
    SDB_S_COMSTRUCT = 3, /* count(|dr| < d)           -> com_struct          dynamics.py:410-428 */
    SDB_S_ROT = 4,     /* sum |dtheta|                -> mean_rotation       dynamics.py:225-227 */
    SDB_S_ROT2 = 5,    /* sum dtheta^2                -> msr                 dynamics.py:229-231 */
    SDB_S_COS = 6,     /* sum cos|dtheta|             -> rot1                dynamics.py:237-247 */
    SDB_S_COS2 = 7,    /* sum (2 cos^2|dtheta| - 1)   -> rot2                dynamics.py:249-259 */
    SDB_S_ROT4 = 8,    /* sum dtheta^4                -> alpha_rot           dynamics.py:261-280 */
    SDB_S_D2R2 = 9,    /* sum |dr|^2 dtheta^2         -> gamma               dynamics.py:282-303 */
    SDB_S_STRUCT = 10, /* count of sites with |site displacement| < d (written by sdb_struct) */
    SDB_S_OVERLAP = 11,/* |top-k(|dr|) & top-k(|dtheta|)| (written by sdb_overlap_*) */
    SDB_S_BADQUAT = 12,/* count of quaternion quotients that are not unit (rowan raises ValueError)
                          or not planar (x or y != 0) */
    SDB_S_OV_RABOVE = 13, /* overlap bookkeeping: |dr|^2 keys above the sampled bracket       */
    SDB_S_OV_TABOVE = 14, /*                      |dtheta| keys above the sampled bracket     */
    SDB_S_OV_BOTH = 15    /*                      molecules above both brackets               */
};

/* One first/last-passage tracker (relaxations.py:187-294,308-335), thresholds pre-digested by the
 * host so that the device compares exactly like numpy compares an f32 array with a Python float:
 *   translation trackers compare on the squared displacement with exact cut points
 *      |dr| >  threshold  <=>  disp2 >= cut_gt     |dr| < threshold  <=>  disp2 < cut_lt
 *   (|dr| = fl32(sqrt(disp2)) is monotone in disp2, the host finds the cut by bisection),
 *   rotation trackers compare |dtheta| directly:  > thr_gt,  < thr_lt,  > irr_gt. */
typedef struct {
    int32_t is_rotation;  /* RelaxationType.rotation / translation          relaxations.py:35-37 */
    int32_t is_last;      /* LastMolecularRelaxation (2 state bits) vs first passage (1 bit) */
    int32_t bit;          /* first bit of this tracker inside the per-molecule flag byte */
    float cut_gt;         /* value >= cut_gt  <=> distance >  threshold        (translation: disp2) */
    float cut_lt;         /* value <  cut_lt  <=> distance <  threshold        (last passage only)  */
    float cut_irr;        /* value >= cut_irr <=> distance >  irreversibility  (last passage only)  */
} SdbTracker;

/* Per-run constants of the fused dynamics step. */
typedef struct {
    int32_t n;            /* molecules */
    int32_t stride;       /* elements per plane of every per-origin / per-frame array; a multiple
                             of 4 that is >= n (16-byte vector accesses) */
    float lx, ly, xy;     /* freud Box(Lx, Ly, xy, is2D=True)                util.py:178-193 */
    float com_cut_lt;     /* disp2 < com_cut_lt <=> |dr| < d, d = pi/(2 wave_number) dynamics.py:399-403;
                             set to -1 when wave_number is None (com_struct stays NaN) */
    int32_t n_trackers;   /* 0..SDB_MAX_TRACKERS; 0 disables the relaxation update */
    int32_t do_sums;      /* 0: skip the reductions (Relaxations.add alone) */
    SdbTracker trackers[SDB_MAX_TRACKERS];
} SdbDynParams;

/* One active time-origin of the current frame (one "keyframe" of read/_read.py:134-141) and the
 * device arrays holding its state.  Planes have `stride` elements (SdbDynParams.stride). */
typedef struct {
    float* d_delta;       /* float[3][stride]: accumulated dx, dy (_util.py:123,149) and rotation
                             about z (_util.py:124,150-153), one plane each */
    uint8_t* d_flags;     /* uint8[stride]: first-passage "fired" bits and last-passage states */
    uint32_t* d_status;   /* uint32[n_trackers][stride]: passage times  relaxations.py:199,218,268-289 */
    uint32_t timediff;    /* frame.timestep - origin.timestep   dynamics.py:213-215, relaxations.py:132 */
    uint32_t flags;       /* SDB_ORIGIN_* */
} SdbOrigin;
#define SDB_ORIGIN_NEW 1u     /* created at this frame: state is initialised instead of read
                                 (Dynamics.from_frame + first compute_all, read/_read.py:135-152) */
#define SDB_ORIGIN_LITERAL 2u /* timediff is not larger than every earlier one of this origin: use
                                 the literal read-modify-write of relaxations.py:215-218 */
#define SDB_ORIGIN_HISTORY 4u /* d_delta is followed by SDB_HISTORY_WORDS 32-bit words (at
                                 d_delta + 3 * stride, zero-initialised by the caller when the origin is
                                 created) in which the `overlap` selection keeps the exact thresholds of
                                 the origin's last two frames; they predict the next bracket, so the
                                 sampling pass is skipped and ~10x fewer candidates are collected */
#define SDB_HISTORY_WORDS 16

/* A run of active origins that share the same previous sample (TrackedMotion.previous_position /
 * previous_orientation, _util.py:155-156).  The minimum-image step and the rotation increment are
 * computed once per segment and molecule and applied to every origin of the segment. */
typedef struct {
    const void* d_prev;   /* float[4][stride] planes (x, y, qw, qz) of the previous sample; NULL means
                             "previous sample == current frame" (origins created at this frame) */
    int32_t first;        /* index of the first origin of the segment in the origin array */
    int32_t count;        /* number of origins in the segment */
    uint32_t flags;       /* SDB_SEG_* */
    uint32_t reserved;    /* 0, or 1 + the index of the increment set precomputed by sdb_dyn_increments
                             for this segment's previous sample */
} SdbSegment;
#define SDB_MAX_INC_SETS 4
#define SDB_SEG_NO_ROTATION 1u /* either sample has no orientation: skip the rotation update
                                  (_util.py:150 "if previous_orientation is not None and ...") */
#define SDB_SEG_ZERO_STEP 2u   /* no new sample: recompute the sums / trackers of the current state
                                  (Dynamics.compute_msd() ... called between add()s); d_pos may be NULL */

/* ---- fused per-frame dynamics + relaxation step (planar: 2D box, quaternions about z) ----------
 * Replaces, for every active origin of one frame, Dynamics.compute_all (dynamics.py:328-394, minus
 * `overlap` and `struct`, see below) and Relaxations.add (relaxations.py:135-162), i.e. the body of
 * the hot loop read/_read.py:134-153.
 *   d_pos  float[n][3], d_quat float[n][4] (may be NULL)  current frame in GSD layout (frame.py:162-184)
 *   d_cur_compact  float[4][stride] out (may be NULL): planes (x, y, qw, qz) of the current frame; it
 *                  becomes the previous sample of every origin updated here
 *   d_origins / d_segments  device copies of the descriptor arrays (n_origins / n_segments entries)
 *   d_partials double[n_origins][sdb_dyn_num_parts(n)][SDB_NSUM] scratch
 *   d_sums  double[n_origins][SDB_NSUM] out, filled by the finalize kernel launched by this call
 *   d_ov_ws workspace of the bracketed `overlap` selection prepared by sdb_overlap_prepare for the
 *           same descriptors (NULL: no overlap bookkeeping), sized for ov_ws_origins origins
 *   d_inc_ws / inc_sets  increments precomputed by sdb_dyn_increments for this frame (NULL / 0: every
 *           segment computes its own); used by the segments whose `reserved` field names a set
 */
int sdb_dyn_step(const SdbDynParams* h_params, const float* d_pos, const float* d_quat,
                 float* d_cur_compact, const SdbOrigin* d_origins, int32_t n_origins,
                 const SdbSegment* d_segments, int32_t n_segments, double* d_partials,
                 double* d_sums, void* d_ov_ws, int32_t ov_ws_origins, void* d_inc_ws,
                 int32_t inc_sets, void* stream);
/* Per-frame increments (minimum-image step and rotation increment of every molecule, the arithmetic of
 * TrackedMotion.add _util.py:149-153) for up to SDB_MAX_INC_SETS previous samples h_prev[i] (device
 * pointers to (x, y, qw, qz) planes; h_seg_flags[i] = SDB_SEG_NO_ROTATION or 0), computed once per
 * frame instead of once per segment.  Also writes d_cur_compact (may be NULL).  d_ws holds
 * sdb_dyn_increments_bytes(n, stride, n_sets) bytes. */
size_t sdb_dyn_increments_bytes(int32_t n, int32_t stride, int32_t n_sets);
int sdb_dyn_increments(const SdbDynParams* h_params, const float* d_pos, const float* d_quat,
                       float* d_cur_compact, const void* const* h_prev, const uint32_t* h_seg_flags,
                       int32_t n_sets, void* d_ws, void* stream);
/* |dr| and |dtheta| of every molecule of one origin (compute_displacement / compute_rotation,
 * dynamics.py:179-181,217-219): d_disp, d_rot float[n] (either may be NULL). */
int sdb_dyn_norms(const float* d_delta, int32_t n, int32_t stride, float* d_disp, float* d_rot,
                  void* stream);
int32_t sdb_dyn_num_parts(int32_t n);

/* Pack a GSD-layout frame into the planar (x, y, qw, qz) form (frame.py:162-184 -> the
 * TrackedMotion.previous_* sample of _util.py:121-131).  Also counts non-planar quaternions into
 * *d_bad (uint32, may be NULL). */
int sdb_pack_frame(const float* d_pos, const float* d_quat, int32_t n, int32_t stride,
                   float* d_compact, uint32_t* d_bad, void* stream);

/* ---- `struct` (dynamics.py:305-326 with _util.py:42-54) ----------------------------------------
 * mode 0 = reference pairing (site-major rotation x molecule-major translation, rotation about x:
 * parity traps T1/T2 of SURVEY.md), mode 1 = physical pairing (in-plane rotation of each molecule's
 * own sites).  Adds the count of sites with displacement < dist into d_sums[o][SDB_S_STRUCT].
 * sites: 3 x (x, y) of the molecule's sites as the f32 values of molecules.py:136-152. */
int sdb_struct(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride,
               const float* h_sites_xy, int32_t n_sites, double dist, int32_t mode, double* d_sums,
               void* stream);

/* ---- `overlap` (dynamics.py:431-445) -----------------------------------------------------------
 * k largest |dr| and k largest |dtheta| per origin, intersection count into
 * d_sums[o][SDB_S_OVERLAP].  Exact; among equal keys the larger molecule indices win.
 *
 * Fused path, three calls per frame around sdb_dyn_step with the same descriptor arrays:
 *   sdb_overlap_prepare  samples the updated keys of every origin and brackets the k-th largest;
 *   sdb_dyn_step         (given d_ov_ws) counts keys above the brackets and lists the candidates;
 *   sdb_overlap_resolve  finds the exact thresholds among the candidates and counts the overlap;
 *                        origins whose bracket failed are redone by the exact kernel below.
 * d_ws: sdb_overlap_ws_bytes(ws_origins, n) bytes, ws_origins >= n_origins. */
size_t sdb_overlap_ws_bytes(int32_t max_origins, int32_t n);
int sdb_overlap_prepare(const SdbDynParams* h_params, const float* d_pos, const float* d_quat,
                        const SdbOrigin* d_origins, int32_t n_origins, const SdbSegment* d_segments,
                        int32_t n_segments, int32_t max_segment_count, int32_t k, void* d_ws,
                        int32_t ws_origins, void* stream);
int sdb_overlap_resolve(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride,
                        int32_t k, void* d_ws, int32_t ws_origins, double* d_sums, void* stream);
/* Diagnostics of the last sdb_overlap_resolve on this workspace (synchronises the stream): h_out[0] =
 * origins that needed the exact kernel, h_out[1 + o] = candidates collected for origin o. */
int sdb_overlap_stats(void* d_ws, int32_t ws_origins, int32_t n, int32_t n_origins, uint32_t* h_out,
                      void* stream);

/* Stand-alone exact selection over the whole arrays (radix select, cooperative kernel).
 * key_mode 0: keys are computed from the (dx, dy, dtheta) planes of each origin's d_delta
 * (|dr|^2 and |dtheta|); key_mode 1: plane 0 holds non-negative displacement keys and plane 2
 * rotation keys (abs applied) -- mobile_overlap() on caller arrays.
 * d_scratch: sdb_overlap_scratch_bytes(n_origins, n) bytes. */
int sdb_overlap(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride,
                int32_t key_mode, int32_t k, void* d_scratch, double* d_sums, void* stream);
size_t sdb_overlap_scratch_bytes(int32_t n_origins, int32_t n);

/* ---- stand-alone trackers on caller-provided distances (API parity for
 * MolecularRelaxation.add relaxations.py:209-218 and LastMolecularRelaxation.add :238-289).
 * dist_is_f64: distances are double (numpy float64 input compares in f64) else float.
 * d_status int64[n] (the reference keeps int64 status words), d_state uint8[n] (last passage). */
int sdb_first_passage(const void* d_dist, int32_t dist_is_f64, int32_t n, double threshold,
                      int64_t timediff, int64_t* d_status, void* stream);
int sdb_last_passage(const void* d_dist, int32_t dist_is_f64, int32_t n, double threshold,
                     double irreversibility, int64_t timediff, int64_t* d_status,
                     uint8_t* d_state, void* stream);

/* ---- general (3D box / arbitrary quaternion) tracked motion, TrackedMotion.add _util.py:133-156.
 * Not the hot path: one origin per call, state in the reference's own (N,3) layout.
 * h_box: Lx, Ly, Lz, xy, xz, yz; is2d as in freud (util.py:178-193).  d_prev_pos / d_prev_quat hold
 * the previous sample and are overwritten with the current one.  d_quat may be NULL (no
 * orientation given); rotate = 1 when both the previous and the current sample have an orientation
 * (_util.py:150).  *d_bad (uint32, may be NULL) counts quotients rowan would reject. */
int sdb_motion_general(const float* h_box, int32_t is2d, int32_t n, const float* d_pos,
                       const float* d_quat, float* d_prev_pos, float* d_prev_quat, int32_t rotate,
                       float* d_dtrans, float* d_drot, uint32_t* d_bad, void* stream);
/* The SDB_S_* sums of Dynamics.compute_* (dynamics.py:179-303) from (N,3) accumulators; optionally
 * also |dr| and |dtheta| per molecule (compute_displacement / compute_rotation).
 * d_partials: double[sdb_sums_general_parts(n)][SDB_NSUM] scratch; d_sums: double[SDB_NSUM]. */
int sdb_sums_general(int32_t n, const float* d_dtrans, const float* d_drot, float com_threshold,
                     float* d_disp, float* d_rot, double* d_partials, double* d_sums, void* stream);
int32_t sdb_sums_general_parts(int32_t n);

/* ---- neighbours / orientational order (order.py:174-309) --------------------------------------
 * k nearest neighbours within rmax (strict rsq < rmax^2, self excluded) in a periodic 2D box,
 * sorted by (rsq, index), padded with 0xFFFFFFFF (nlist) / -1 (rsq).
 * Outputs (any may be NULL): d_nlist uint32[n][k], d_rsq float[n][k], d_counts uint32[n],
 * d_order double[n] = mean_k cos^2(2*intrinsic_distance) with missing neighbours -> self
 * (order.py:54-65,85-86; needs d_quat float[n][4]).  k <= SDB_MAX_K. */
#define SDB_MAX_K 16
int sdb_neighbours(const float* h_box, int32_t n, const float* d_pos, const float* d_quat,
                   float rmax, int32_t k, uint32_t* d_nlist, float* d_rsq, uint32_t* d_counts,
                   double* d_order, void* d_scratch, size_t scratch_bytes, void* stream);
size_t sdb_neighbours_scratch_bytes(const float* h_box, int32_t n, float rmax);
/* The same search with capacity k (d_counts saturates at k, like num_neighbours' k = 9, order.py:279-281) and rows
 * of width k_list <= k in d_nlist / d_rsq / the order parameter: ONE build serves compute_neighbours (k_list = 8),
 * orientational_order (8) and num_neighbours (k = 9) of a frame, where the reference builds three times. */
int sdb_neighbours_ex(const float* h_box, int32_t n, const float* d_pos, const float* d_quat, float rmax, int32_t k,
                      int32_t k_list, uint32_t* d_nlist, float* d_rsq, uint32_t* d_counts, double* d_order,
                      void* d_scratch, size_t scratch_bytes, void* stream);
/* order._orientational_order (order.py:68-86) and order._neighbour_relative_angle (order.py:28-65)
 * on a caller-provided neighbour list: d_order double[n] and/or d_angles double[n][k]. */
int sdb_orientational_order(const uint32_t* d_nlist, int32_t n, int32_t k, const float* d_quat,
                            double* d_order, double* d_angles, void* stream);

/* ---- misc ---------------------------------------------------------------------------------------*/
/* ---- Voronoi neighbour counts ---------------------------------------------------------------------
 * freud.voronoi.Voronoi(box, buff).computeNeighbors(position).nlist.neighbor_counts of
 * order.compute_voronoi_neighs (order.py:312-315) for an orthorhombic 2D box: the cell of every molecule is
 * clipped by the bisectors of its neighbours in order of distance (d_nlist: (n, k) from sdb_neighbours with
 * radius rmax, k <= SDB_MAX_K) until no farther molecule can cut it; molecules whose list ends before that
 * are redone against all molecules.  d_counts: uint32[n] (0xFFFFFFFF: the cell meets the molecule's own
 * periodic images); d_scratch: uint32[n + 1]. */
int sdb_voronoi_counts(const float* h_box, int32_t n, const float* d_pos, const uint32_t* d_nlist, int32_t k,
                       float rmax, uint32_t* d_counts, uint32_t* d_scratch, void* stream);

/* ---- radial distribution function (wave-number estimate) -------------------------------------------
 * Pair-distance histogram of freud.density.RDF(rmax, dr).compute(box, positions), the heavy part of
 * calculate_wave_number (dynamics/_util.py:66-86), which Relaxations runs when it is given no wave number
 * (relaxations.py:61-62): every ordered pair i != j with rsq = |box.wrap(r_j - r_i)|^2 < rmax^2 (f32, freud
 * operation order) adds one count to bin floorf(sqrtf(rsq) / dr), bins >= nbins are dropped.
 * h_box: Lx, Ly, Lz, xy, xz, yz; d_pos: float[n][3]; d_counts: uint64[nbins] (overwritten); nbins <= 1024.
 * All pairs are visited (the reference asks for rmax = min(L)/2.2): O(n^2). */
int sdb_rdf_histogram(const float* h_box, int32_t is2d, int32_t n, const float* d_pos, float rmax, float dr,
                      int32_t nbins, uint64_t* d_counts, void* stream);

/* Page-lock / release a host range the caller keeps mapped (e.g. the mmap of a GSD trajectory, frame.py:162-184's
 * arrays): frames inside it can be passed to sdb_engine_step as SDB_MEM_PINNED.  Failure (SDB_ECUDA) leaves no
 * pending CUDA error: the caller simply keeps using SDB_MEM_HOST. */
int sdb_host_register(void* h_ptr, size_t bytes);
int sdb_host_unregister(void* h_ptr);

/* cudaMemcpyAsync(host -> device) on `stream`: the upload of a frame slice by a host that keeps its own device
 * buffers (the multi-GPU sliced exchange) without going through a tensor library's synchronising copy. */
int sdb_upload_async(void* d_dst, const void* h_src, size_t bytes, void* stream);
/* Diagnostics: the engine's multi-threaded host copy (SDB_MEM_HOST staging) on caller buffers. */
void sdb_test_host_copy(void* h_dst, const void* h_src, size_t bytes);

const char* sdb_last_error(void);
const char* sdb_version(void);
/* Number of kernels launched by this library since load (used by bench.py's gpu_launches). */
uint64_t sdb_launch_count(void);
/* Host-side evaluation of the device arithmetic helpers (same source, compiled for the host) so
 * the CPU test-suite can check them against the oracle without a GPU.  Test hooks only. */
double sdb_test_rot_increment(float w, float z, float pw, float pz, int32_t* bad);
void sdb_test_wrap2d(float lx, float ly, float xy, float vx, float vy, float* ox, float* oy);

/* ---- intermediate scattering function -----------------------------------------------------------
 * Dynamics.compute_isf (dynamics.py:233-235; wave vectors :105-112): for every origin the SUM over
 * molecules of the mean over `angular_resolution` wave vectors of length `wave_number` of
 * cos(k . dr); divide by n for the reference's value.  d_out double[n_origins] (cleared by the call).
 * sdb_isf takes one (x, y) array pair whose components are elem_stride floats apart per molecule
 * ((N,3) row-major displacement array: d_dx = base, d_dy = base + 1, elem_stride = 3). */
int sdb_isf_origins(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride,
                    double wave_number, int32_t angular_resolution, double* d_out, void* stream);
int sdb_isf(const float* d_dx, const float* d_dy, int32_t elem_stride, int32_t n, double wave_number,
            int32_t angular_resolution, double* d_out, void* stream);

/* ---- one call per frame -------------------------------------------------------------------------
 * The body of the reference hot loop (read/_read.py:134-153) for every active origin, enqueued from
 * native code: upload of the descriptor block (h_desc, pinned: n_origins SdbOrigin records, then at
 * byte segments_offset the n_segments SdbSegment records; desc_bytes in total) to d_desc, then
 * sdb_overlap_prepare -> sdb_dyn_step -> sdb_struct -> sdb_overlap_resolve with the same meaning of
 * the arguments as those calls (overlap_k <= 0 or d_ov_ws NULL: no `overlap`; h_sites_xy NULL: no
 * `struct`), then the download of d_sums[n_origins][SDB_NSUM] to h_sums (pinned; NULL: stay on the
 * device).  inc_sets > 0: sdb_dyn_increments(h_inc_prev, h_inc_flags, inc_sets, d_inc_ws) runs first and
 * the segments whose `reserved` field names a set use its increments.  Asynchronous on `stream`.
 * The overlap bookkeeping and the row sums (small, latency-bound kernels) run on an internal
 * high-priority side stream next to the increments pass and the struct kernel and are joined before
 * the download (SDB_SIDE_STREAM=0: everything on `stream`).
 * step_flags: SDB_STEP_PREDICTED -- every origin has taken part in the two previous frames (increasing
 * time, sums requested), so its overlap bracket is predicted from its threshold history and the sample
 * kernels are skipped; an origin for which that does not hold takes the exact selection kernel. */
#define SDB_STEP_PREDICTED 1
int sdb_frame_step(const SdbDynParams* h_params, const float* d_pos, const float* d_quat,
                   float* d_cur_compact, const void* h_desc, void* d_desc, int32_t desc_bytes,
                   int32_t segments_offset, int32_t n_origins, int32_t n_segments,
                   int32_t max_segment_count, double* d_partials, double* d_sums, double* h_sums,
                   void* d_ov_ws, int32_t ov_ws_origins, int32_t overlap_k, const float* h_sites_xy,
                   int32_t n_sites, double struct_dist, int32_t struct_mode, const void* const* h_inc_prev,
                   const uint32_t* h_inc_flags, int32_t inc_sets, void* d_inc_ws, int32_t step_flags,
                   void* stream);

/* Per-kernel-group device times of sdb_frame_step (CUDA events on the launching stream), for
 * bench.py's roofline: groups 0 overlap prepare (main-stream part), 1 increments pass + dyn_step, 2 struct,
 * 3 what is left of the side stream's row sums + overlap resolve after struct.
 * sdb_profile_collect waits for the recorded events, returns the summed times in ms (h_ms[4]), the
 * number of steps and of origin-steps, and clears the collection. */
void sdb_profile_enable(int32_t on);
int sdb_profile_collect(double* h_ms, int64_t* h_steps, int64_t* h_origin_steps);

/* ---- multi-GPU frame exchange (origin-sharded path; the reference is single-process) -------------
 * Copy `bytes` (a multiple of 16) from local device memory to an NVSwitch multicast address with
 * multimem.st stores: one pass on the reader rank delivers the frame to the same offset of every
 * rank's symmetric buffer.  d_multicast_dst is the multicast virtual address of the destination
 * (torch symmetric memory: handle.multicast_ptr + offset).  blocks <= 0 picks a default grid. */
int sdb_multicast_copy(const void* d_src, void* d_multicast_dst, size_t bytes, int32_t blocks, void* stream);

/* ==== native engine: the whole hot loop behind one call per frame =====================================
 * Replaces the body of reference read/_read.py:128-168 *including its bookkeeping*: the dict of per-keyframe
 * (Dynamics, Relaxations) objects (_read.py:117,135-141), TrackedMotion.previous_* (_util.py:155-156, here
 * reference-counted planar frames), the grouping of the active keyframes by previous sample, the descriptor
 * rings, the staging of a host frame through pinned memory (frame.py:162-184 -> device) and the asynchronous
 * download of the result rows.  A host in any language calls
 *     sdb_engine_step(engine, &frame, stream)  ->  ticket          (asynchronous, ~10 CUDA calls inside)
 *     sdb_engine_wait(engine, ticket, rows, ...)                   (blocks until that frame's rows are on the host)
 * once per frame with HOST or DEVICE buffers; no torch types, no Python between the kernels.
 * The engine owns its device memory (cudaMalloc): per-origin state, previous-sample pool, rings, workspaces.
 * One engine = one device = one host thread at a time. */
typedef struct SdbEngine SdbEngine;

#define SDB_MEM_DEVICE 0 /* frame pointers are device pointers (16-byte aligned), used in place */
#define SDB_MEM_HOST 1   /* pageable host memory (numpy array, memory-mapped trajectory file): copied into the
                            engine's pinned ring by a small pool of host threads, then uploaded on a copy stream */
#define SDB_MEM_PINNED 2 /* page-locked host memory: uploaded on the copy stream directly; the caller keeps it
                            unchanged until the step has completed */

#define SDB_ENGINE_DRY 1u        /* test hook: no CUDA calls at all (bookkeeping and descriptors only) */
#define SDB_ENGINE_NO_PREDICT 2u /* sampled overlap brackets every frame (no SDB_ORIGIN_HISTORY) */
#define SDB_ENGINE_NO_INCREMENTS 4u /* never run the per-frame increments pass */

typedef struct {
    int32_t n;                /* molecules */
    float box[6];             /* Lx, Ly, Lz, xy, xz, yz of frame.py:190-196 (2D: Lz, xz, yz unused) */
    double wave_number;       /* <= 0: none (com_struct, struct stay unset: dynamics.py:369-372,385-390) */
    float com_cut_lt;         /* disp2 < com_cut_lt <=> |dr| < pi / (2 wave_number) in f32 (see SdbDynParams) */
    int32_t n_trackers;       /* pre-digested trackers (see SdbTracker); 0: Dynamics only */
    SdbTracker trackers[SDB_MAX_TRACKERS];
    int32_t compute_sums;     /* 0: Relaxations.add only */
    int32_t overlap_k;        /* int(0.1 N) of dynamics.py:441; 0: no `overlap` */
    int32_t n_sites;          /* sites of the molecule for `struct` (molecules.py:136-152); 0: no `struct` */
    float sites_xy[16];
    int32_t struct_mode;      /* 0 reference pairing, 1 physical (see sdb_struct) */
    int32_t compute_isf;      /* != 0: scattering_function column (dynamics.py:364-366) */
    int32_t angular_resolution;
    int32_t ring_depth;       /* frames in flight (>= 2) */
    int32_t inc_max_per_segment; /* <= 0: default 40 (see sdb_dyn_increments) */
    int32_t chunk;            /* > 0: fixed number of origins per segment (tests); 0: automatic */
    uint32_t flags;           /* SDB_ENGINE_* */
} SdbEngineConfig;

#define SDB_FRAME_NO_SUMS 1u     /* Relaxations.add only for this frame (no rows) */
#define SDB_FRAME_ZERO_STEP 2u   /* no new sample: recompute the rows of the current state (pos/quat ignored) */
#define SDB_FRAME_ROWS_DEVICE 4u /* rows stay on the device (d_rows_out and/or sdb_engine_device_rows) */

typedef struct {
    int64_t timestep;         /* frame.timestep */
    const float* pos;         /* float[n][3] */
    const float* quat;        /* float[n][4] (w, x, y, z) or NULL */
    int32_t where;            /* SDB_MEM_* */
    uint32_t flags;           /* SDB_FRAME_* */
    const int64_t* keys;      /* the active keyframes (`indexes` of read/_read.py:128); unknown keys are
                                 created at this frame (Dynamics.from_frame, _read.py:135-141) */
    int32_t n_keys;
    int32_t reserved;
    double* d_rows_out;       /* optional DEVICE destination of the n_keys x SDB_NSUM rows, in `keys` order
                                 (may be peer memory of another GPU: multi-GPU row collection) */
} SdbFrame;

int sdb_engine_create(const SdbEngineConfig* config, SdbEngine** out);
void sdb_engine_destroy(SdbEngine* engine);
/* Relaxations.set_mol_relax (relaxations.py:108-133): new tracker set, every origin's passage state afresh. */
int sdb_engine_set_trackers(SdbEngine* engine, const SdbTracker* trackers, int32_t n_trackers);
/* One frame.  Returns a ticket >= 0, or a negative SDB_E* code. */
int64_t sdb_engine_step(SdbEngine* engine, const SdbFrame* frame, void* stream);
/* Wait for a step and copy its rows out, in the order of the step's `keys`: h_rows double[n_keys][SDB_NSUM]
 * (NULL for steps without host rows), h_timediffs int64[n_keys], h_isf double[n_keys] (any may be NULL).
 * A ticket stays valid for ring_depth - 1 further steps (SDB_EINVAL afterwards). */
int sdb_engine_wait(SdbEngine* engine, int64_t ticket, double* h_rows, int64_t* h_timediffs, double* h_isf);
/* Device copy of the rows of a step (valid like the ticket, and until a step with more active origins than ever
 * before re-sizes the rings), in the engine's internal origin order;
 * h_perm int32[n_keys] (may be NULL): position of keys[i] in that order. */
int sdb_engine_device_rows(SdbEngine* engine, int64_t ticket, double** d_rows, int32_t* h_perm);
/* A whole batch of frames in one call (the frame loop itself in native code): steps frames[0..n_frames) in
 * order, keeping ring_depth - 1 of them in flight, and writes the rows of frame f at
 * h_rows + SDB_NSUM * (sum of n_keys of the frames before f), likewise h_timediffs / h_isf (may be NULL).
 * Frames flagged NO_SUMS or ROWS_DEVICE leave their rows untouched. */
int sdb_engine_run(SdbEngine* engine, const SdbFrame* frames, int32_t n_frames, double* h_rows, int64_t* h_timediffs,
                   double* h_isf, void* stream);
/* Drop a keyframe and recycle its state. */
int sdb_engine_drop(SdbEngine* engine, int64_t key);

typedef struct {
    int64_t timestep;         /* of the origin's first frame */
    int64_t updates;          /* frames the origin has been stepped with, the one that created it included */
    int64_t max_timediff;
    float* d_delta;           /* float[3][stride] (+ SDB_HISTORY_WORDS words) */
    uint8_t* d_flags;         /* uint8[stride] or NULL */
    uint32_t* d_status;       /* uint32[n_trackers][stride] or NULL */
    const float* d_prev;      /* float[4][stride] planes of the previous sample */
    int32_t prev_has_orientation;
    int32_t stride;
} SdbOriginInfo;
int sdb_engine_origin(SdbEngine* engine, int64_t key, SdbOriginInfo* out); /* SDB_EINVAL: unknown key */
int32_t sdb_engine_num_origins(SdbEngine* engine);
int32_t sdb_engine_keys(SdbEngine* engine, int64_t* h_keys, int32_t capacity); /* in creation order */

typedef struct {
    int64_t frames, updates;          /* frames stepped, molecule x origin updates performed */
    int64_t h2d_bytes, d2h_bytes;     /* of the last step */
    int64_t pool_frames, pool_free;   /* previous-sample pool */
    int64_t device_bytes;             /* device memory owned by the engine */
    int32_t last_segments, last_inc_sets, last_step_flags, capacity;
    void* d_ov_ws;                    /* overlap workspace (for sdb_overlap_stats) */
    void* d_inc_ws;
} SdbEngineStats;
int sdb_engine_stats(SdbEngine* engine, SdbEngineStats* out);
/* Test hook (dry engines): the descriptor block of the last step -- n_origins SdbOrigin records followed, at
 * byte *segments_offset, by n_segments SdbSegment records.  Returns the number of bytes copied. */
int32_t sdb_engine_last_descriptors(SdbEngine* engine, void* h_out, int32_t capacity, int32_t* n_origins,
                                    int32_t* n_segments, int32_t* segments_offset);

#ifdef __cplusplus
}
#endif
#endif /* SDB200_H */
