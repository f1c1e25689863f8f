/* Test infrastructure: a plain-C client of include/sdb200.h.  It proves that the header is valid C99 on its own
 * (no torch, no C++), reports the layout a C compiler gives every struct of the boundary (the test compares it
 * with the ctypes mirrors of sdanalysis_b200/_lib.py), and drives a dry engine through the exported entry points
 * the way a maintainer's own binding would: the frame loop of read/_read.py:128-168 for a linear schedule. */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sdb200.h"

#define SIZE(T) printf("size %s %zu\n", #T, sizeof(T))
#define FIELD(T, f) printf("field %s %s %zu\n", #T, #f, offsetof(T, f))

static void layouts(void) {
    SIZE(SdbTracker); FIELD(SdbTracker, is_rotation); FIELD(SdbTracker, is_last); FIELD(SdbTracker, bit);
    FIELD(SdbTracker, cut_gt); FIELD(SdbTracker, cut_lt); FIELD(SdbTracker, cut_irr);
    SIZE(SdbDynParams); FIELD(SdbDynParams, n); FIELD(SdbDynParams, stride); FIELD(SdbDynParams, lx);
    FIELD(SdbDynParams, ly); FIELD(SdbDynParams, xy); FIELD(SdbDynParams, com_cut_lt);
    FIELD(SdbDynParams, n_trackers); FIELD(SdbDynParams, do_sums); FIELD(SdbDynParams, trackers);
    SIZE(SdbOrigin); FIELD(SdbOrigin, d_delta); FIELD(SdbOrigin, d_flags); FIELD(SdbOrigin, d_status);
    FIELD(SdbOrigin, timediff); FIELD(SdbOrigin, flags);
    SIZE(SdbSegment); FIELD(SdbSegment, d_prev); FIELD(SdbSegment, first); FIELD(SdbSegment, count);
    FIELD(SdbSegment, flags); FIELD(SdbSegment, reserved);
    SIZE(SdbEngineConfig); FIELD(SdbEngineConfig, n); FIELD(SdbEngineConfig, box);
    FIELD(SdbEngineConfig, wave_number); FIELD(SdbEngineConfig, com_cut_lt); FIELD(SdbEngineConfig, n_trackers);
    FIELD(SdbEngineConfig, trackers); FIELD(SdbEngineConfig, compute_sums); FIELD(SdbEngineConfig, overlap_k);
    FIELD(SdbEngineConfig, n_sites); FIELD(SdbEngineConfig, sites_xy); FIELD(SdbEngineConfig, struct_mode);
    FIELD(SdbEngineConfig, compute_isf); FIELD(SdbEngineConfig, angular_resolution);
    FIELD(SdbEngineConfig, ring_depth); FIELD(SdbEngineConfig, inc_max_per_segment); FIELD(SdbEngineConfig, chunk);
    FIELD(SdbEngineConfig, flags);
    SIZE(SdbFrame); FIELD(SdbFrame, timestep); FIELD(SdbFrame, pos); FIELD(SdbFrame, quat); FIELD(SdbFrame, where);
    FIELD(SdbFrame, flags); FIELD(SdbFrame, keys); FIELD(SdbFrame, n_keys); FIELD(SdbFrame, reserved);
    FIELD(SdbFrame, d_rows_out);
    SIZE(SdbOriginInfo); FIELD(SdbOriginInfo, timestep); FIELD(SdbOriginInfo, updates);
    FIELD(SdbOriginInfo, max_timediff); FIELD(SdbOriginInfo, d_delta); FIELD(SdbOriginInfo, d_flags);
    FIELD(SdbOriginInfo, d_status); FIELD(SdbOriginInfo, d_prev); FIELD(SdbOriginInfo, prev_has_orientation);
    FIELD(SdbOriginInfo, stride);
    SIZE(SdbEngineStats); FIELD(SdbEngineStats, frames); FIELD(SdbEngineStats, updates);
    FIELD(SdbEngineStats, h2d_bytes); FIELD(SdbEngineStats, d2h_bytes); FIELD(SdbEngineStats, pool_frames);
    FIELD(SdbEngineStats, pool_free); FIELD(SdbEngineStats, device_bytes); FIELD(SdbEngineStats, last_segments);
    FIELD(SdbEngineStats, last_inc_sets); FIELD(SdbEngineStats, last_step_flags); FIELD(SdbEngineStats, capacity);
    FIELD(SdbEngineStats, d_ov_ws); FIELD(SdbEngineStats, d_inc_ws);
    printf("const SDB_NSUM %d\nconst SDB_MAX_TRACKERS %d\nconst SDB_OK %d\nconst SDB_EINVAL %d\nconst SDB_ECUDA %d\n",
           SDB_NSUM, SDB_MAX_TRACKERS, SDB_OK, SDB_EINVAL, SDB_ECUDA);
    printf("const SDB_MEM_HOST %d\nconst SDB_MEM_PINNED %d\nconst SDB_ENGINE_DRY %u\nconst SDB_FRAME_NO_SUMS %u\n",
           SDB_MEM_HOST, SDB_MEM_PINNED, SDB_ENGINE_DRY, SDB_FRAME_NO_SUMS);
}

#define MUST(call)                                                                                     \
    do {                                                                                               \
        long long rc_ = (long long)(call);                                                             \
        if (rc_ < 0) {                                                                                 \
            printf("fail %s -> %lld: %s\n", #call, rc_, sdb_last_error());                             \
            return 1;                                                                                  \
        }                                                                                              \
    } while (0)

/* Linear schedule: a new origin every `gap` steps, at most `live` of them alive (the oldest is dropped). */
static int dry_loop(int n, int n_frames, int gap, int live) {
    SdbEngineConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n = n;
    cfg.box[0] = cfg.box[1] = 70.0f;
    cfg.box[2] = 1.0f;
    cfg.wave_number = 2.9;
    cfg.com_cut_lt = 0.2933f;
    cfg.compute_sums = 1;
    cfg.overlap_k = n / 10;
    cfg.ring_depth = 3;
    cfg.flags = SDB_ENGINE_DRY;
    SdbEngine* eng = NULL;
    MUST(sdb_engine_create(&cfg, &eng));
    float* pos = (float*)calloc((size_t)n * 3, sizeof(float));
    float* quat = (float*)calloc((size_t)n * 4, sizeof(float));
    for (int i = 0; i < n; ++i) quat[4 * i] = 1.0f;
    int64_t keys[64], diffs[64];
    int64_t first_key = 0;
    for (int f = 0; f < n_frames; ++f) {
        int64_t newest = f / gap;
        if (newest - first_key + 1 > live) {
            MUST(sdb_engine_drop(eng, first_key));
            ++first_key;
        }
        int n_keys = 0;
        for (int64_t k = first_key; k <= newest; ++k) keys[n_keys++] = k;
        SdbFrame frame;
        memset(&frame, 0, sizeof frame);
        frame.timestep = f;
        frame.pos = pos;
        frame.quat = quat;
        frame.where = SDB_MEM_HOST;
        frame.keys = keys;
        frame.n_keys = n_keys;
        int64_t ticket = sdb_engine_step(eng, &frame, NULL);
        MUST(ticket);
        MUST(sdb_engine_wait(eng, ticket, NULL, diffs, NULL));
        printf("frame %d keys %d diffs", f, n_keys);
        for (int i = 0; i < n_keys; ++i) printf(" %lld", (long long)diffs[i]);
        printf("\n");
    }
    SdbOriginInfo info;
    MUST(sdb_engine_origin(eng, first_key, &info));
    printf("oldest key %lld timestep %lld updates %lld max_timediff %lld stride %d\n", (long long)first_key,
           (long long)info.timestep, (long long)info.updates, (long long)info.max_timediff, info.stride);
    printf("unknown key rc %d\n", sdb_engine_origin(eng, 1000, &info));
    SdbEngineStats st;
    MUST(sdb_engine_stats(eng, &st));
    printf("stats frames %lld updates %lld origins %d\n", (long long)st.frames, (long long)st.updates,
           sdb_engine_num_origins(eng));
    sdb_engine_destroy(eng);
    free(pos);
    free(quat);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "layout") == 0) {
        layouts();
        return 0;
    }
    if (argc > 1 && strcmp(argv[1], "dry") == 0) {
        printf("version %s\n", sdb_version());
        return dry_loop(1000, 12, 3, 3);
    }
    fprintf(stderr, "usage: cabi_client layout|dry\n");
    return 2;
}
