// engine.h — internal definitions of librts_b200 (device layouts, engine state, kernel launchers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/rts_b200.h"

#ifndef RTS_QNODES
#define RTS_QNODES 0            // 1: later waves walk the quantised 32-byte nodes (measured slower: 1.61 vs 1.37 ms; kept as a tuning build)
#endif
#define RTS_LEAF_MAX 2          // triangles per BVH leaf (collapsed LBVH subtrees); measured best of 1/2/4 on B200
#define RTS_STACK_DEPTH 96      // traversal stack entries per thread
#ifndef RTS_WAVE_BLOCK
#define RTS_WAVE_BLOCK 128      // threads per CTA of the bounce-wave kernel
#endif
#ifndef RTS_WAVE_MIN_BLOCKS
#define RTS_WAVE_MIN_BLOCKS 6           // resident CTAs per SM the register allocation must allow (later waves)
#endif
#ifndef RTS_WAVE_MIN_BLOCKS_PRIMARY
#define RTS_WAVE_MIN_BLOCKS_PRIMARY 8   // same, primary wave (its state is smaller)
#endif
#define RTS_MAX_RX 64
#define RTS_SMEM_BINS 256u        // dense bin tables up to this many bins are pre-reduced per CTA in shared memory (12 KB)
#define RTS_COOP_STACK 4096u     // entries of a warp's shared stack in the cooperative traversal of a straggler ray (follow.cuh)
#define RTS_MAX_WORLD 16          // GPUs of one node that can share a peer-memory bin exchange (comm.cu)
#define RTS_EAGER_BINS 256u      // bins brought to pinned host memory right behind a pulse
#define RTS_REBUILD_RATIO 1.2   // refit falls back to a rebuild when SAH cost exceeds this x the as-built cost

// ---- device layouts -------------------------------------------------------------------------
// BVH node, 64 B = 3 x 128-bit + 1 x 64-bit loads: the two child boxes and two child references.
// A box is stored as centre c and half-extent h (fp32) with [c-h, c+h] containing the fp32 box that the
// reference's bound program produces (triangle_mesh.cu:228-229, rounded outward) — h is rounded up.
// The words are ordered so that each 64-bit half of a 128-bit load is one operand of a packed fp32
// instruction (FFMA2 / FADD2 on sm_100a): (x,y) pairs per child, (child0, child1) pairs for z.
// child ref >= 0 : index of an internal node;  < 0 : leaf, ~ref = (first_leaf_pos << 3) | (count-1).
struct __align__(64) BvhNode {
    float c0x, c0y, h0x, h0y;   // child 0: centre xy, half-extent xy
    float c1x, c1y, h1x, h1y;   // child 1
    float c0z, c1z, h0z, h1z;   // z of both children
    int32_t ref0, ref1;
    int32_t pad[2];
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

// Quantised BVH node, 32 B = 2 x 128-bit loads, for every ray that starts inside the scene (all later waves): the
// traversal loop is bound by the L1 data pipe (128 B per clock per SM; a warp of 32 lanes fetching 64-byte nodes keeps
// it busy for 16 clocks per visit, twice the time the visit's instructions take to issue), so the node is halved.
// Each plane of the two child boxes is a 15-bit position on a frame around the scene: x(q) = lo + q * ext / 32768,
// stored as h = 0x8000 | q so that one PRMT with 0x3F800000 yields the float 1 + q/32768 (k_pack; traverse_q in
// trace.cu holds the error analysis).  Boxes are rounded outward and widened by one more cell.
struct __align__(32) QNode {
    uint32_t w[6];              // child 0: x, y, z; child 1: x, y, z — low half = lower plane, high half = upper plane
    int32_t ref0, ref1;         // as in BvhNode
};
static_assert(sizeof(QNode) == 32, "QNode must be 32 bytes");
struct QFrame { float lo[3], ext[3]; uint32_t overflow; uint32_t pad; };   // device-resident; overflow: a box left the frame

// Leaf-ordered triangle record, 80 B = 5 x 128-bit loads: world-space fp64 vertices + ids.
struct __align__(16) TriRec {
    double p[9];            // p0.xyz p1.xyz p2.xyz
    uint32_t tri_id;        // global triangle id (target offset + local index)
    uint32_t target;
};
static_assert(sizeof(TriRec) == 80, "TriRec must be 80 bytes");

// Wavefront ray state, struct of arrays (one array per field).
#define RTS_NF 14
enum RayField { F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_LEN, F_PW, F_DOP, F_FX, F_FY, F_FZ, F_N0, F_N1 };
struct RayQueue {
    double *f[RTS_NF];
    unsigned long long *key;   // packed target-path row so far (digits base n_targets+1)
    uint32_t *ray;             // primary ray index (global launch index)
    uint32_t *meta;            // reflDepth[0:8) refrDepth[8:10) slot[10:14) end[14] primary[15] col[16:20)
};

// control block and work items of the projected primary wave (raster.cuh)
struct RasterCtl {
    unsigned long long area;   // summed footprint (candidate ray/triangle pairs) of this batch
    unsigned n_items;          // row chunks of large footprints
    unsigned next_item;        // work counter of k_raster_big (chunks are taken dynamically: their sizes differ a lot)
};
struct RasterItem { unsigned pos, z0, z1, pad; };

struct Counters {              // device-side, accumulated with atomics
    unsigned long long segments, hits, shaded, captured, multi, edge, refracted, nodes, tris, overflow, kept;
};

struct RxDev { double cx, cy, cz, radius, min_theta, max_theta, min_phi, max_phi; };

// Everything a bounce-wave kernel needs, passed by value.
// device mirrors of rts_table2d / rts_antenna (values: device pointer)
struct DevTable { uint32_t n_az, n_el; double az0, az_step, el0, el_step; const double *values; };
struct DevAntenna { DevTable gain; double bore_az, bore_el, rate_az, rate_el; double pos[3]; };

struct WaveParams {
    // scene
    const BvhNode *nodes;
    const QNode *qnodes;            // the same tree, quantised (later waves)
    const QFrame *qframe;
    const TriRec *trirec;
    int32_t root_ref;
    uint32_t n_tris;
    const float *scene_abs;         // [3] device: max |coordinate| of the scene box per axis (slab error bound)
    const double *world_normals;    // [Nn*3]
    const uint32_t *tris;           // [T*3] local vertex indices
    const uint32_t *t_norm_off;     // per target
    const uint32_t *t_tri_off;
    const uint32_t *t_per_face;     // per target: 1 when n_normals > n_verts
    const double *t_refl, *t_refr;  // per target
    const double *t_vel;            // per target, 3
    // launch constants (ray_tracer.cu:144-204, hoisted on the host)
    uint32_t nx, ny, nz;
    uint64_t R3;
    uint32_t dMax, rMax, D, M, W;
    uint32_t interpolate;
    double origin[3];
    double beamStart[3];
    double slope[3];
    double Rot[9], Rot1[9];
    double boresight[3];
    int single_ray;
    // receivers
    uint32_t n_rx;
    const RxDev *rx;
    // post-process constants (ray_tracer.cpp:1233-1253, aggregation.cu:59-60)
    double cspeed, carrier, wl2gain;
    const double *t_rcs;            // per-target scalar RCS (Target::GetRCS stand-in, ray_tracer.cpp:1226) or NULL = 1
    uint64_t B;                     // path-key base = n_targets + 1
    // path key
    uint64_t powB[RTS_MAX_DEPTH + 1];
    uint64_t key_all;               // sum_{c<D} B^c
    // shard
    uint64_t ray_begin, ray_stride, n_primary;   // primary rays of this batch: index = ray_begin + (batch_base + i)*ray_stride
    uint64_t batch_base;            // first shard-local primary index of this batch
    // primary-ray tiling: shard-local indices below swz_limit are visited in 8x4 tiles of the (y,z) launch
    // plane (one tile per warp) instead of 32x1 strips; swz_w = row length in local indices, 0 = off
    uint32_t swz_w;
    uint64_t swz_limit;
    // queues
    RayQueue in, out;
    const unsigned long long *in_count;
    unsigned long long *out_count;
    // With refraction the queue is filled from both ends: reflected rays from slot 0 upwards (in_count / out_count),
    // refracted children from slot capacity-1 downwards (in_back / out_back), so that a warp of the next wave holds one
    // kind of ray.  Entry i of the wave is slot i for i < *in_count, else slot capacity-1-(i-*in_count).  nullptr: one end.
    const unsigned long long *in_back;
    unsigned long long *out_back;
    unsigned long long out_capacity;
    unsigned long long *work_counter;
    // outputs
    uint32_t flags;
    double *bin_sums;               // [n_bins*5]
    unsigned long long *bin_mins;   // [n_bins]
    uint64_t n_bins;
    // sparse bins (when n_rx * (K+1)^D is too large for a dense table): open-addressing table over the same sums / mins
    // arrays, keyed by rx * (K+1)^D + path key; hash_used lists the occupied slots (emission and clearing cost what is occupied)
    unsigned long long *hash_keys;  // [n_bins] (n_bins = table size, a power of two); ~0 = empty; nullptr: dense table
    uint32_t *hash_used;            // [n_bins]
    unsigned *hash_count;
    rts_ray_record *results;        // records mode
    int32_t *targ_intersect;
    double *rcs_angle;
    int32_t *tri_path;
    Counters *counters;
    unsigned long long *wave_segs;  // [32] segments traced per wave index
    uint32_t wave_index;
    uint32_t chain_below;           // later waves with fewer queued rays than this follow reflections in place
    // primary visibility by projection (raster.cuh); raster_ctl == nullptr: off
    double AT[9];                   // transpose of Rot1*Rot: world offset from the Tx -> beam frame
    double *dirs[3];                // [batch] primary ray directions (struct of arrays)
    unsigned long long *hits;       // [batch] (fp32 t bits << 32 | triangle id), ~0 = none
    RasterCtl *raster_ctl;
    RasterItem *raster_items;
    uint32_t raster_item_cap;
    uint32_t raster_chunk;          // candidates per row chunk of a large footprint (k_raster_big: one warp per chunk)
    const uint32_t *leaf_of_tri;
    uint32_t hits_resolved;         // k_raster_resolve ran: hit words carry leaf positions (and the kept flag), not triangle ids
    // the shard's columns as a lattice of the image (nx == 1, stride divides ny): y = lat_c0 + k * stride,
    // shard-local index = z * lat_w + k - lat_g0; lat_w == 0: no lattice (general begin / stride)
    uint32_t lat_w, lat_c0;
    uint64_t lat_g0;
    // footprint pass over a subset: triangle ids of the moving targets (raster_list), or everything but them (raster_skip:
    // per-target flags); raster_static: control block of the cached static pass, counted by the guard
    const uint32_t *raster_list;
    uint32_t raster_list_count;
    const uint32_t *raster_skip;
    const RasterCtl *raster_static;
    // kept first-reflection hits (coherent.cuh); w1_static == nullptr: off
    unsigned long long *w1_static;          // [batch] closest static hit of each pixel's first reflection (t bits << 32 | leaf pos), ~0 = none
    const unsigned long long *hits_static;  // [batch] kept primary hits (t bits << 32 | triangle id)
    const uint32_t *moving_flags;           // per target
    const BvhNode *mover_nodes;             // bounding boxes of the moving targets, two per node
    uint32_t n_mover_nodes;
    unsigned long long *fill_counter;       // work counter of k_wave1_fill
    uint32_t *todo_list;                    // queue indices k_wave1_kept leaves to the ordinary wave kernel (nullptr: all)
    unsigned long long *todo_count;
    // split later waves (split.cuh): k_traverse answers the closest-hit query of every queued ray into trav_hits[slot]
    // (fp32 t bits << 32 | leaf position; TRAV_MISS = no hit but a receiver sphere is crossed; TRAV_DEAD = nothing left
    // to do), k_shade_wave then shades / captures the survivors.  split_below: waves with fewer rays stay with the fused
    // kernel (which returns at once for the others when split_on is set).
    unsigned long long *trav_hits;
    uint32_t split_on, split_below, split_keep_all;
    // per-CTA shared-memory copy of a small dense bin table (trace.cu: bins_smem_*): > 0 = the number of bins cached,
    // i.e. n_bins; the kernels are then launched with n_bins * 48 bytes of dynamic shared memory
    uint32_t smem_bins;
    // tabulated callbacks (RTS_TABLES): per-target RCS tables (nullptr: scalar t_rcs at capture), antennas (nullptr: scalar
    // gains inside wl2gain); keep_first: the first hit point travels with the ray state (gains need it, like records mode)
    const DevTable *rcs_tab;
    const DevAntenna *ant_tx, *ant_rx;
    uint32_t keep_first;
    double wl2, gain_tx_scalar, gain_rx_scalar;   // Wl^2 (ray_tracer.cpp:1247) and the scalar gains, apart, when antennas are in use
    // follow.cuh: per-warp scratch stacks of the warp-cooperative traversal that finishes straggler rays
    // (RTS_COOP_STACK entries per resident warp of k_primary_follow)
    int *coop_stacks;
};

// ---- engine ---------------------------------------------------------------------------------
// Pinned host block the device writes results of the pulse in flight into (asynchronous read-back).
struct Readback {
    Counters counters;
    unsigned long long wave_segs[32];
    double sah;
    RasterCtl raster, raster_static;
    uint32_t bins_count;
    unsigned long long comm_timed_out;   // peer-memory exchange: a rank never arrived (comm.cu)
};

// One slot of the pinned staging ring for small per-pulse host arrays (poses, receivers, velocities, RCS).
struct StageSlot {
    char *host = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    bool in_flight = false;
};

// What the primary ray directions of a batch depend on (raster.cuh: k_primary_dirs); equal key = reusable buffer.
struct DirsKey {
    uint32_t nx, ny, nz;
    int single;
    uint64_t begin, stride, base, n;
    double c[30];   // origin, beamStart, slope, Rot, Rot1, boresight
};

// Peer-memory exchange of the receiver bins (comm.cu): this rank's exchange block and the mapped blocks of its peers.
struct Comm {
    char *block = nullptr;
    uint64_t bytes = 0, max_bins = 0, half_words = 0;
    uint32_t rank = 0, world = 0;
    char *peer[RTS_MAX_WORLD] = {};
    bool ipc_opened[RTS_MAX_WORLD] = {};
    bool connected = false, check_timeout = false;
    unsigned long long seq = 0;
    int clock_khz = 1965000;
};

// Tuning / test switches: read from the environment once in rts_create, changed afterwards only through rts_set_option.
struct Knobs {
    int no_chain = 0, no_raster = 0, no_tiles = 0, one_ended_queue = 0, debug_raster = 0, no_static_hits = 0,
        no_kept_reflections = 0, no_split = 1, no_follow = 0, no_smem_bins = 0, no_overlap = 0, debug_timeline = 0, no_split_raster = 0;
    long long batch = 0;           // 0 = default 2^24 primaries per batch
    int hash_bins = 0;             // 1: sparse (hashed) bins also where a dense table would fit
    uint32_t hash_log2 = 22;       // slots of the sparse bin table
    uint32_t split_below = 1u << 18;   // later waves with at least this many rays run as k_traverse + k_shade_wave
};

struct rts_engine {
    int device = 0;
    Knobs knobs;
    int num_sms = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    // Work of a pulse that does not depend on the rest of the stream runs beside it (knob no_overlap turns both off):
    //   side_dirs  (lowest stream priority) k_primary_dirs needs the launch geometry only and two buffers whose last readers
    //              are the previous batch's / pulse's shading pass and the BVH primary wave behind it (ev_dirs_free); in a
    //              from-scratch pulse with moving targets the footprints of the triangles that never move follow it on this
    //              stream (their records are not touched by the pose update), otherwise the engine's stream waits for the
    //              directions (ev_dirs_done) and projects everything itself.  So this work runs beside the previous pulse's
    //              last waves and bin emission and this pulse's pose update.  (Measured and not kept, DESIGN §6: a second
    //              set of buffers so that the direction pass can run beside the previous pulse's footprint kernels or
    //              shading pass — those need the fp64 pipe / the issue slots as much as it does, the sum stays the same.)
    //   side_bvh   the refit above the moving triangles (k_fit, k_pack, scene_abs, SAH read-back) forks behind the leaf
    //              boxes / triangle records (ev_bvh_fork); only traversal needs the nodes, so the footprint kernels run
    //              beside it and the first kernel that walks the tree waits for ev_bvh_done (bvh_join).
    cudaStream_t side_dirs = nullptr, side_bvh = nullptr;
    cudaEvent_t ev_dirs_free = nullptr, ev_dirs_done = nullptr, ev_bvh_fork = nullptr, ev_bvh_done = nullptr;
    bool dirs_free_valid = false, bvh_join_pending = false;
    bool split_static = false;     // this pulse projected the never-moving triangles on side_dirs (its own RasterCtl: d_raster_ctl_static)
    // knob debug_timeline: timestamps of up to 16 pulses' direction pass / footprint kernels / shading pass, printed by rts_sync
    cudaEvent_t tl_ev[16][7] = {};   // + start and end of the later waves
    int tl_n = 0;
    cudaEvent_t ev[6] = {};
    cudaEvent_t wave_ev[34] = {};
    cudaEvent_t split_ev[3] = {};      // around k_traverse and k_shade_wave of the second wave
    cudaEvent_t follow_ev[2] = {};     // around k_primary_follow
    bool follow_timed = false;
    float follow_ms = 0.f;
    bool split_timed = false;
    bool followed = false;             // this pulse's primary shading pass traced the first reflections in place (follow.cuh)
    int follow_grid = 0;
    float split_ms[2] = {};
    unsigned long long *d_wave_segs = nullptr;
    float wave_ms[32] = {};
    unsigned long long wave_segs[32] = {};
    uint32_t n_waves = 0;
    uint64_t launches = 0;

    // host-side scene meta
    uint32_t n_targets = 0, n_tris = 0, n_verts = 0, n_normals = 0;
    std::vector<uint32_t> tri_off, vert_off, norm_off, t_nverts, t_ntris, t_nnormals;
    bool scene_ready = false;

    // device scene
    double *d_base_verts = nullptr, *d_base_normals = nullptr, *d_world_verts = nullptr, *d_world_normals = nullptr;
    uint32_t *d_tris = nullptr, *d_tri_target = nullptr, *d_vert_target = nullptr, *d_norm_target = nullptr;
    uint32_t *d_t_vert_off = nullptr, *d_t_norm_off = nullptr, *d_t_tri_off = nullptr, *d_t_per_face = nullptr;
    double *d_t_refl = nullptr, *d_t_refr = nullptr, *d_t_vel = nullptr, *d_t_rcs = nullptr;
    rts_pose *d_poses = nullptr;

    // BVH
    unsigned long long *d_morton = nullptr, *d_morton_sorted = nullptr;
    uint32_t *d_order_in = nullptr, *d_order = nullptr, *d_leaf_of_tri = nullptr;
    float *d_tri_box = nullptr, *d_node_box = nullptr, *d_scene_box = nullptr;
    int32_t *d_parent = nullptr;     // [2T-1]: internal i at i, leaf pos p at (T-1)+p
    int2 *d_children = nullptr, *d_range = nullptr;
    uint32_t *d_fit_flags = nullptr;
    BvhNode *d_nodes = nullptr;
    QNode *d_qnodes = nullptr;
    QFrame *d_qframe = nullptr;
    TriRec *d_trirec = nullptr;
    void *d_cub_temp = nullptr;
    size_t cub_temp_bytes = 0;
    int32_t root_ref = 0;
    int leaf_max = RTS_LEAF_MAX;       // 1..8; RTS_LEAF_MAX env var overrides (tuning)
    int builder_forced = 0;            // RTS_BVH=lbvh (1) | ploc (2); 0 = choose per scene by SAH cost
    int builder = 0;                   // the scene's topology builder: 0 undecided, 1 Morton radix tree, 2 PLOC (ploc.cuh)
    rts_bvh_info bvh_info = {};
    float *d_scene_abs = nullptr;      // [3] max |coordinate| of the scene box per axis, refreshed by every update
    // partial refit: only the targets that have moved since the scene was committed
    std::vector<rts_pose> h_poses;     // poses last applied
    std::vector<uint8_t> moving;       // per target: pose has differed from the committed one at least once
    bool partial_ready = false;
    uint32_t *d_moving = nullptr, *d_vlist = nullptr, *d_nlist = nullptr, *d_tlist = nullptr, *d_nodelist = nullptr;
    uint32_t *d_list_counts = nullptr, *d_mark = nullptr;
    float *d_static_box = nullptr;     // ordered-uint encoded, like d_scene_box
    double *d_sah_static = nullptr;
    uint32_t n_dv = 0, n_dn = 0, n_dt = 0, n_dnode = 0;
    bool sah_pending = false, refit_timed = false;
    uint32_t refits_since_build = 0;
    cudaEvent_t sah_ev = nullptr;
    unsigned long long *d_violations = nullptr;
    double *d_sah = nullptr;
    double sah_at_build = 0;
    uint32_t builds = 0;

    // wave state
    RayQueue q[2] = {};
    void *q_slab[2] = {nullptr, nullptr};
    uint64_t q_capacity = 0;
    unsigned long long *d_counts = nullptr;   // [96] queue counts (front), work counters, queue counts (back)
    Counters *d_counters = nullptr;
    RxDev *d_rx = nullptr;
    int wave_grid = 0, wave_grid_primary = 0, trav_grid = 0;

    // primary visibility by projection
    double *d_dirs = nullptr;
    unsigned long long *d_hits = nullptr;
    void *d_raster_ctl = nullptr, *d_raster_items = nullptr;
    uint64_t raster_alloc = 0;
    DirsKey dirs_key = {};
    bool dirs_valid = false;
    // closest hits of the triangles that never move, kept while launch geometry, scene and moving set stay the same
    unsigned long long *d_hits_static = nullptr;
    void *d_raster_ctl_static = nullptr;
    bool static_valid = false;
    uint64_t scene_version = 0, moving_version = 0, static_scene_version = 0, static_moving_version = 0;
    // kept first-reflection hits (coherent.cuh)
    unsigned long long *d_w1_static = nullptr;
    unsigned *d_target_box = nullptr;
    BvhNode *d_mover_nodes = nullptr;
    uint32_t *d_todo = nullptr;
    char *d_coop_stacks = nullptr;     // follow.cuh: COOP_WARP_BYTES of scratch per resident warp of the wave kernels
    int coop_ctas = 0;
    uint64_t todo_alloc = 0;
    unsigned long long *d_trav_hits = nullptr;   // split.cuh: one hit word per queue slot
    uint64_t trav_alloc = 0;
    bool coh_on = false, coh_fill = false, w1_valid = false;
    uint32_t w1_builds = 0, w1_interp = 0, w1_dmax = 0, w1_rmax = 0;

    // outputs
    double *d_bin_sums = nullptr;
    unsigned long long *d_bin_mins = nullptr;
    uint64_t n_bins_dense = 0, bins_alloc = 0;   // n_bins_dense: entries of the table in use (dense bins, or slots of the hash table)
    bool bins_hashed = false, hash_ready = false;
    unsigned long long *d_hash_keys = nullptr;
    uint32_t *d_hash_used = nullptr;
    unsigned *d_hash_count = nullptr;
    uint64_t hash_alloc = 0, bins_per_rx = 0;
    unsigned long long *d_ckeys = nullptr, *d_cmins = nullptr;   // compact copies of the occupied bins (multi-GPU exchange)
    double *d_csums = nullptr;
    uint64_t compact_alloc = 0;
    rts_bin *d_bins_out = nullptr, *h_bins = nullptr;   // h_bins: pinned; the current one of h_bins_buf
    bool bins_eager = false;
    // Two pinned blocks, used in turn by the pulses whose bins are brought to the host right behind them: a host that has
    // enqueued pulse p+1 (RTS_ASYNC) can still read pulse p's bins (rts_get_bins_previous) — one wait for an event that
    // was recorded behind pulse p's copies, while pulse p+1 runs.
    rts_bin *h_bins_buf[2] = {nullptr, nullptr};
    uint32_t *h_bins_count = nullptr;                   // pinned, [2]
    cudaEvent_t bins_ev[2] = {nullptr, nullptr};
    int bins_slot = 0, prev_bins_slot = 0;
    bool prev_bins_eager = false;
    bool emit_precleared = false;      // the receiver totals / emitted-bin count were zeroed by the pulse's clear kernel
    double *d_rx_sums = nullptr;
    unsigned long long *d_rx_mins = nullptr;
    uint32_t *d_bins_out_count = nullptr;
    uint64_t bins_out_alloc = 0;
    rts_ray_record *d_results = nullptr;
    int32_t *d_targ_intersect = nullptr, *d_tri_path = nullptr;
    double *d_rcs_angle = nullptr;
    uint64_t rec_alloc_rays = 0, rec_alloc_D = 0, rec_alloc_W = 0;

    // asynchronous plumbing
    StageSlot stage[8];
    int stage_next = 0;
    Readback *h_rb = nullptr;          // pinned
    bool pulse_pending = false;        // a pulse was enqueued and its read-back not yet folded into `stats`
    bool pulse_single_batch = false, pulse_raster = false;
    uint64_t pulse_primary = 0, pulse_waves = 0;

    Comm comm;
    // tabulated callbacks (rts_set_rcs_tables / rts_set_antennas)
    DevTable *d_rcs_tab = nullptr;
    double *d_rcs_values = nullptr;
    uint32_t n_rcs_tab = 0;
    DevAntenna *d_ant_tx = nullptr, *d_ant_rx = nullptr;
    double *d_ant_values = nullptr;
    uint32_t n_ant_rx = 0;

    // last pulse
    bool have_pulse = false, bins_finalised = false;
    uint32_t last_flags = 0;
    rts_sizes last_sizes = {};
    uint32_t last_B = 1, last_D = 0, last_nrx = 0;
    uint64_t last_begin = 0, last_stride = 1, last_n_primary = 0;   // the shard of the last pulse
    rts_stats stats = {};
};

// error plumbing (api.cu)
int rts_fail(int code, const char *fmt, ...);
// NVTX ranges around the host-side stages (tracing hooks); without a profiler attached a range is one
// function-pointer test.  Kernels of a stage show up under its range on the nsys / ncu timeline.
#include <nvtx3/nvToolsExt.h>
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

#define RTS_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return rts_fail(RTS_ERR_CUDA, "%s failed: %s (%s:%d)", #call,            \
                                               cudaGetErrorString(_e), __FILE__, __LINE__);             \
    } while (0)

// bvh.cu
int bvh_alloc(rts_engine *e);
void bvh_free(rts_engine *e);
int bvh_update_world(rts_engine *e);          // transform + tri boxes (+ tri records when topology exists)
int bvh_build(rts_engine *e);                 // full LBVH build at current world geometry
int bvh_refit(rts_engine *e);                 // bottom-up refit + repack
int bvh_check(rts_engine *e, uint64_t *violations);
void bvh_join(rts_engine *e);                // the engine's stream waits for a refit still running on side_bvh
int bvh_read_scene_box(rts_engine *e);        // fills bvh_info.scene_lo/hi (synchronises)
void bvh_sync_info(rts_engine *e);            // waits for the last refit's SAH cost

// api.cu: pinned staging for host arrays that are copied to the device asynchronously
char *stage_acquire(rts_engine *e, size_t bytes);   // pinned pointer valid until stage_release
void stage_release(rts_engine *e);                  // call after enqueuing the copies that read it
int pulse_collect(rts_engine *e);                   // fold the read-back of an asynchronous pulse into e->stats (synchronises)

// trace.cu
int trace_alloc_queues(rts_engine *e, uint64_t capacity);
int trace_launch_wave(rts_engine *e, const WaveParams &p, bool primary, bool records);
int trace_launch_split(rts_engine *e, WaveParams &p, bool records);   // k_traverse + k_shade_wave ahead of the fused kernel
int trace_raster_alloc(rts_engine *e, uint64_t batch);     // buffers of the projected primary wave
int trace_launch_raster(rts_engine *e, WaveParams &p, bool records, bool single_batch);   // enqueue it (before the BVH primary wave)
int trace_launch_kept(rts_engine *e, WaveParams &p, bool records);   // second wave: rays served from the kept first-reflection hits
int trace_wave_grid(rts_engine *e);

// aggregate.cu
int agg_collect_bins(rts_engine *e, rts_bin *out, uint32_t cap, uint32_t *n);
int agg_hash_prepare(rts_engine *e, uint64_t slots);   // allocate / clear the sparse bin table for a pulse
int agg_hash_compact(rts_engine *e, void **keys, void **sums, void **mins, uint32_t *n);
int agg_hash_load(rts_engine *e, const void *keys, const void *sums, const void *mins, uint32_t n);
int agg_emit_bins_async(rts_engine *e);
int agg_collect_bins_previous(rts_engine *e, rts_bin *out, uint32_t cap, uint32_t *n);
int agg_pulse_clear(rts_engine *e, bool dense_bins, uint64_t n_bins, uint32_t n_rx);   // one launch for everything a pulse starts from
int agg_fill_records(rts_engine *e, uint64_t ray_total, uint32_t D, uint32_t W);
int agg_get_received(rts_engine *e, uint64_t cap, uint64_t *n, uint64_t *slots, rts_ray_record *results, int32_t *targ_intersect,
                     double *rcs_angle);
int agg_get_records_shard(rts_engine *e, uint64_t *n_shard, rts_ray_record *results, int32_t *targ_intersect, double *rcs_angle,
                          int32_t *tri_path);
int agg_kernel_wrapper(rts_engine *e, rts_ray_record *rx_results, const int32_t *rx_intersects, uint32_t received,
                       uint32_t depth_total, double cspeed, double carrier, double *npath, double *power,
                       double *doppler, double *delay, double *phase, int32_t *path_match);
