/* rts_b200.h — C-ABI of librts_b200.so, the B200-native (sm_100a) ray-tracing radar path.
 *
 * Drop-in boundary for the reference's OptiX pipeline.  Each entry point names the reference
 * interface it replaces (file:line in /root/reference).  All pointers passed in are caller-owned
 * HOST memory unless a name ends in _device; the library owns every device allocation behind the
 * opaque handle.  Every function returns RTS_OK (0) or a negative error code and never exits the
 * process (the reference aborts: RT_CHECK_ERROR, aggregation.cu:17-27, ray_tracer.cpp:455-458);
 * rts_last_error() returns the thread-local message of the last failure.
 *
 * There is no CPU fallback: every compute entry point fails with RTS_ERR_NO_DEVICE when no
 * sm_100-class CUDA device is usable.
 */
#ifndef RTS_B200_H
#define RTS_B200_H

#include "rts_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RTS_OK               0
#define RTS_ERR_ARG         -1
#define RTS_ERR_NO_DEVICE   -2
#define RTS_ERR_CUDA        -3
#define RTS_ERR_STATE       -4
#define RTS_ERR_CAPACITY    -5

/* rts_trace_pulse flags */
#define RTS_OUT_BINS      1u  /* fused post-process + aggregation into (receiver, path) bins         */
#define RTS_OUT_RECORDS   2u  /* reference-shaped per-ray arrays (dbuf_results / _targ_intersect / _rcs_angle) + tri_path */
#define RTS_COUNT_NODES   4u  /* fill rts_stats.nodes_visited / tris_tested (slower)                 */
#define RTS_NO_FINALISE   8u  /* leave bins un-finalised (caller reduces across GPUs, then rts_finalise_bins) */
#define RTS_NO_RCS_ANGLES 16u /* records mode: skip the four atan2 per bounce, leave rcs_angle at -1e6 */
#define RTS_ASYNC         32u /* return once the pulse is enqueued on the engine's stream; rts_sync / any getter waits */
#define RTS_TABLES        128u /* fused bins: per-hop RCS from the tables of rts_set_rcs_tables and gains from the antennas of
                                 rts_set_antennas (angle-dependent Target::GetRCS / GetGain evaluated on the device) */
#define RTS_NO_REUSE      64u /* trace this pulse from scratch: use nothing kept from earlier pulses (ray directions, static
                                 primary hits, static first-reflection hits — raster.cuh, coherent.cuh)                 */

typedef struct rts_engine rts_engine;

/* Rigid pose of one target for one pulse: world = (has_rotation ? R * base : base) + t, evaluated
 * exactly as ray_tracer.cpp:156-170 + :1006-1014 do ((0 + R[i][0]*v0) + R[i][1]*v1) + R[i][2]*v2, then + t[i]).
 * `base` is the mesh given to rts_scene_set_targets (already carrying the t = 0 rotation,
 * ray_tracer.cpp:956-987).  Normals get the rotation only. */
typedef struct rts_pose {
    double  R[9];          /* row-major */
    double  t[3];
    int32_t has_rotation;
    int32_t _pad;
} rts_pose;

typedef struct rts_sizes {
    uint64_t rays;        /* nx*ny*nz                                        */
    uint64_t ray_total;   /* M * rays      (ray_tracer.cpp:608-626)          */
    uint32_t depth_total; /* D = max_refl + (max_refr ? 2 : 0)   (:655)      */
    uint32_t slots;       /* M                                               */
    uint32_t tri_cols;    /* W = max_refl + 3, columns of tri_path           */
    uint32_t _pad;
} rts_sizes;

typedef struct rts_bvh_info {
    uint32_t n_tris, n_nodes, root_is_leaf, max_leaf;
    float    scene_lo[3], scene_hi[3];
    float    ms_build, ms_refit;
    double   sah_cost;     /* sum of internal-node box areas of the current tree (m^2)                 */
    double   sah_at_build; /* the same when the topology was last built; refit rebuilds beyond 1.2 x   */
    uint32_t builds;       /* number of full builds so far                                             */
    uint32_t builder;      /* topology of the current tree: 1 = Morton radix tree (LBVH), 2 = PLOC     */
} rts_bvh_info;

/* ---- life cycle (replaces rtContextCreate / rtContextDestroy, ray_tracer.cpp:532-534,1358) ---- */
int         rts_create(int device, rts_engine **out);
void        rts_destroy(rts_engine *e);
const char *rts_last_error(void);
const char *rts_version(void);
/* sizeof() of the POD structs as compiled, for binding self-checks:
 * [0] rts_ray_record [1] rts_target_mesh [2] rts_rx_sphere [3] rts_rx_desc [4] rts_pulse
 * [5] rts_bin [6] rts_stats [7] rts_pose [8] rts_response [9] rts_sizes */
int         rts_abi_sizes(uint32_t sizes[10]);
/* Run subsequent work of this engine on a caller-provided cudaStream_t (NULL = engine's own).  The engine keeps two side
 * streams of its own, ordered against this one by events (the direction pass of the projected primary wave, at the lowest
 * stream priority, and the BVH refit above the moving triangles); a caller's stream created above the lowest priority
 * (cudaStreamCreateWithPriority) lets the direction pass of the next pulse fill what the last waves of the previous pulse
 * leave instead of competing with them.  Waits for the engine's earlier work. */
int         rts_set_stream(rts_engine *e, void *cuda_stream);
/* Tuning / test switches (no counterpart in the reference, whose only knobs are MaxThreads / MaxBlocks,
 * ray_tracer.cpp:512).  The environment variable RTS_<NAME> gives an option its initial value when the engine is created;
 * the environment is never read again afterwards.  Names: "bvh" (0 = choose by SAH cost, 1 = Morton radix tree, 2 = PLOC),
 * "leaf_max" (1..8), "no_chain", "no_raster", "no_tiles", "one_ended_queue", "debug_raster", "no_static_hits",
 * "no_kept_reflections", "no_split", "split_below" (rays), "no_follow", "no_smem_bins", "no_split_raster", "no_overlap" (1 = no
 * side streams: every kernel of a pulse on the engine's stream), "debug_timeline" (rts_sync prints when the direction pass,
 * the footprint kernels, the shading pass and the later waves of the last pulses ran), "batch" (primaries per batch, 0 = 2^24),
 * "hash_bins" (1 = sparse bin table also where a dense one would fit), "hash_log2" (log2 of its slots, default 22). */
int         rts_set_option(rts_engine *e, const char *name, int64_t value);

/* ---- host helpers (pure host code, usable without a GPU) ---- */
/* Receiver sphere centre and angular window: ray_tracer.cpp:894-918. */
void rts_rx_sphere_from_desc(const rts_rx_desc *desc, rts_rx_sphere *out);
/* Result-array sizes: ray_tracer.cpp:600-626, 655. */
int  rts_result_sizes(const rts_pulse *pulse, rts_sizes *out);
/* Mesh generators and rigid rotation: ray_tracer.cpp:156-170, 226-297, 300-426, 429-504.
 * Two-call protocol: NULL output arrays → counts only. */
int  rts_rect_mesh(float w, float h, float d, float yaw, float pitch, float roll,
                   double *verts, uint32_t *n_verts, uint32_t *tris, uint32_t *n_tris,
                   double *normals, uint32_t *n_normals);
int  rts_sphere_mesh(uint32_t subdivs, float radius, float yaw, float pitch, float roll,
                     double *verts, uint32_t *n_verts, uint32_t *tris, uint32_t *n_tris,
                     double *normals, uint32_t *n_normals);
int  rts_file_mesh(const char *v_file, const char *n_file, float yaw, float pitch, float roll,
                   double *verts, uint32_t *n_verts, uint32_t *tris, uint32_t *n_tris,
                   double *normals, uint32_t *n_normals);
/* Rotation matrix Rz(yaw)*Ry(pitch)*Rx(roll) with the reference's float angles (ray_tracer.cpp:156-162). */
void rts_rotation_matrix(float yaw, float pitch, float roll, double R[9]);

/* ---- scene (replaces the per-pulse rtGeometry/rtBuffer/rtAcceleration set-up,
 *      ray_tracer.cpp:1017-1133, and OptiX's "Bvh" builder) ---- */
/* Upload all targets in their base pose and build the BVH (Morton LBVH) on the device. */
int rts_scene_set_targets(rts_engine *e, const rts_target_mesh *targets, uint32_t n_targets);
/* Per-pulse target motion (ray_tracer.cpp:936-1014): transform on the device, refit the BVH; falls back to
 * a full rebuild when the refitted tree's SAH cost has drifted more than 20 % above its as-built cost. */
int rts_scene_set_poses(rts_engine *e, const rts_pose *poses, uint32_t n_targets);
/* Full rebuild at the current poses (what the reference does every pulse, ray_tracer.cpp:1126-1130). */
int rts_scene_rebuild(rts_engine *e);
int rts_scene_bvh_info(rts_engine *e, rts_bvh_info *out);
/* Parity/diagnostic read-backs: world-space vertices of one target [n_verts*3]; leaf boxes in
 * global-triangle order [n_tris*6] (the `bound` program, triangle_mesh.cu:204-233). */
int rts_scene_get_world_vertices(rts_engine *e, uint32_t target, double *out);
int rts_scene_get_tri_bounds(rts_engine *e, float *out6);
/* Returns 0 when every triangle's box is contained in all its ancestors' boxes, else the number of violations. */
int rts_scene_check_bvh(rts_engine *e, uint64_t *violations);

/* ---- one pulse (replaces rtContextLaunch3D + the result hand-off, ray_tracer.cpp:1165-1258) ---- */
int rts_trace_pulse(rts_engine *e, const rts_pulse *pulse, uint32_t flags);
/* Angle-dependent callbacks as data (ray_tracer.cpp:1219-1247 without the per-ray host loop).
 * rts_set_rcs_tables: one table per target, sampled by the caller from Target::GetRCS(az, el, Wl) over the summed
 * in/out angles the reference stores in dbuf_rcs_angle (normal_shader.cu:259-265, 320-326: az in [-2pi, 2pi], el in
 * [-pi, pi]); a target whose table has n_az == 0 keeps its scalar rts_pulse.targ_rcs (or 1).  The factor of a hop is
 * folded into the ray's power where the reference writes that hop's angles.  Pulses without refraction only
 * (max_refr == 0): the reference's pre-filled path rows of refracted chains are not reproduced by the fused form.
 * rts_set_antennas: transmitter and per-receiver gain patterns + boresights + positions (rts_antenna); NULL keeps the
 * scalar rts_pulse.gain_tx / gain_rx.  The tables are copied to the device; n_targets == 0 / NULL clears.
 * A pulse uses them when traced with RTS_OUT_BINS | RTS_TABLES. */
int rts_set_rcs_tables(rts_engine *e, const rts_table2d *tables, uint32_t n_targets);
int rts_set_antennas(rts_engine *e, const rts_antenna *tx, const rts_antenna *rx, uint32_t n_rx);
/* Wait for everything enqueued on the engine's stream (RTS_ASYNC pulses, pose updates) and report a deferred error. */
int rts_sync(rts_engine *e);
int rts_get_stats(rts_engine *e, rts_stats *out);
/* Per-bounce-wave profile of the last pulse: device milliseconds (CUDA events on the engine's stream) and
 * the number of ray segments each wave traced, summed over ray batches.  *n = number of waves. */
int rts_get_wave_profile(rts_engine *e, uint32_t cap, float *ms, uint64_t *segments, uint32_t *n);
/* Device milliseconds of the second wave's two kernels in the last pulse (split.cuh): ms[0] = k_traverse (closest-hit
 * queries of every first reflection), ms[1] = k_shade_wave (closest_hit / miss / bins of the survivors); both 0 when
 * that wave ran as the fused kernel.  The wave-level figure of rts_get_wave_profile also holds kernels that returned at once. */
int rts_get_split_profile(rts_engine *e, float ms[2]);
/* Device milliseconds of k_primary_follow in the last pulse (follow.cuh: the projected primary wave's shading pass with
 * every first reflection traced in place — the longest single kernel of a from-scratch pulse); 0 when the pulse did not
 * run it (kept first-reflection hits in use, BVH primary wave, max_refl == 0, option no_follow, several batches). */
int rts_get_follow_profile(rts_engine *e, float *ms);
/* Cumulative number of CUDA kernels this engine has launched (every <<<>>> of the library). */
int rts_kernel_launches(rts_engine *e, uint64_t *out);
/* Measurement aid (bench.py roofline): read bandwidth in GB/s of a `bytes`-sized device buffer streamed
 * `reps` times with 128-bit loads that bypass L1 — L2 bandwidth when the buffer fits the 126 MB L2, HBM bandwidth
 * when it is several times larger.  Blocks until the measurement is done. */
int rts_probe_read_bandwidth(rts_engine *e, uint64_t bytes, uint32_t reps, double *gbs);
/* RTS_OUT_BINS: non-empty bins sorted by (rx, path). *n is the total even when cap is smaller. */
int rts_get_bins(rts_engine *e, rts_bin *out, uint32_t cap, uint32_t *n);
/* Pipelined host loops: the bins of the pulse BEFORE the last one enqueued.  A host that traces with RTS_ASYNC can enqueue
 * pulse p+1 and then collect pulse p here — it waits only for pulse p's own read-back (two pinned blocks are used in turn),
 * so the GPU never idles while the host consumes a pulse; pulses are independent (ray_tracer.cpp:843), the order of the
 * responses a simulator receives is unchanged.  Needs pulse p traced with RTS_OUT_BINS, finalised (by the pulse itself,
 * rts_finalise_bins or rts_comm_allreduce_bins) and at most 256 non-empty bins; RTS_ERR_STATE / RTS_ERR_CAPACITY otherwise.
 * Counters and wave profile (rts_get_stats ...) always describe the last pulse. */
int rts_get_bins_previous(rts_engine *e, rts_bin *out, uint32_t cap, uint32_t *n);
/* RTS_OUT_BINS: the responses the reference would emit for this pulse (ray_tracer.cpp:1289-1320): one per
 * unique d_pathMatch value, i.e. one per non-direct bin plus one for a receiver's direct bin when a direct
 * ray is that receiver's first received ray; sorted by representative slot like sort+unique(:1291-1292).
 * The host forms InterpPoint(power, t + delay, delay, doppler, phase, noise_temperature(rx)) from each.
 * *n is the total even when cap is smaller. */
int rts_get_responses(rts_engine *e, rts_response *out, uint32_t cap, uint32_t *n);
/* RTS_OUT_RECORDS: copy out the reference-shaped arrays; any pointer may be NULL.
 *   results [ray_total] · targ_intersect [ray_total*D] · rcs_angle [ray_total*D*2] · tri_path [ray_total*W] */
int rts_get_records(rts_engine *e, rts_ray_record *results, int32_t *targ_intersect, double *rcs_angle,
                    int32_t *tri_path);

/* RTS_OUT_RECORDS: the same arrays for the rays of the last pulse's shard only (pulse->ray_begin / ray_count / ray_stride: one
 * rank's share of a large launch, or a strided sample), gathered on the device.  *n_shard = rays in the shard; ray k of the
 * shard (launch index ray_begin + k*ray_stride) has its result slot s at index k + s*n_shard:
 *   results [M*n_shard] · targ_intersect [M*n_shard*D] · rcs_angle [M*n_shard*D*2] · tri_path [M*n_shard*W]
 * Call with every array NULL for the count; any array may be NULL. */
int rts_get_records_shard(rts_engine *e, uint64_t *n_shard, rts_ray_record *results, int32_t *targ_intersect,
                          double *rcs_angle, int32_t *tri_path);

/* RTS_OUT_RECORDS: only the received rays (received >= 0), compacted on the device in result-slot order — the
 * h_rx_results / h_rx_intersects arrays the reference's host loop collects (ray_tracer.cpp:1190-1221) before it applies
 * the Target::GetRCS / GetGain callbacks (:1226-1247), plus each ray's result-slot index and its rcs_angle row.
 * Call with cap = 0 for the count; then results[cap], targ_intersect[cap*D], rcs_angle[cap*D*2], slots[cap] (any may be
 * NULL).  Feed the scaled records to rts_aggregate / rs::kernel_wrapper. */
int rts_get_received(rts_engine *e, uint64_t cap, uint64_t *n, uint64_t *slots, rts_ray_record *results,
                     int32_t *targ_intersect, double *rcs_angle);

/* ---- multi-GPU plumbing: raw bin accumulators for an external reduction (NCCL all-reduce) ----
 * sums_device: double[n_bins_dense*5] {npath, Σ√P, Σdelay, Σphase, ΣDoppler}  (reduce with SUM)
 * mins_device: uint64[n_bins_dense]   smallest result-slot index; empty bins hold 0x7f7f7f7f7f7f7f7f, so a MIN
 *              reduction over either uint64 or int64 views is correct */
int rts_bins_device(rts_engine *e, void **sums_device, uint64_t *n_sum_doubles, void **mins_device,
                    uint64_t *n_mins);
/* The same for pulses whose bins live in the sparse table (more than 2^20 dense bins, or option "hash_bins"; rts_bins_device
 * then fails with RTS_ERR_STATE).  rts_bins_compact_device gathers this GPU's occupied bins into compact device arrays —
 * keys uint64[n] (rx * (K+1)^D + path key), sums double[n*5], mins uint64[n] — and returns n (it waits for the pulse).  The
 * ranks exchange them (all-gather of keys -> sorted union -> SUM / MIN all-reduce over the union, rts_b200/dist.py) and hand
 * the merged arrays (distinct keys) back with rts_bins_load_compact; then rts_finalise_bins as for the dense table.
 * Replaces nothing in the reference, which has no multi-GPU path; it groups arbitrary path rows (aggregation.cu:43-57). */
int rts_bins_compact_device(rts_engine *e, void **keys_device, void **sums_device, void **mins_device, uint32_t *n);
int rts_bins_load_compact(rts_engine *e, const void *keys_device, const void *sums_device, const void *mins_device, uint32_t n);
/* Marks the (externally reduced) bins final and enqueues their emission on the engine's stream: call it after the
 * reduction has been enqueued on that same stream (rts_set_stream) or has completed. */
int rts_finalise_bins(rts_engine *e);
/* ---- Peer-memory exchange of the receiver bins (one engine per GPU of one NVLink/NVSwitch node; comm.cu) -----------
 * Replaces the pair of NCCL all-reduces behind a ray- or pulse-sharded launch (the reference has no multi-GPU path; the
 * exchanged quantity is the commutative aggregation of aggregation.cu:56-69).  Every rank creates an exchange block,
 * the ranks swap its handle (rts_comm_ipc_handle between processes — any transport, e.g. an all-gather of 64 bytes —
 * or rts_comm_local_ptr between the threads of one process after cudaDeviceEnablePeerAccess) and connect; from then on
 * rts_comm_allreduce_bins, called by every rank once per pulse traced with RTS_OUT_BINS | RTS_NO_FINALISE, enqueues on
 * the engine's stream: publish this rank's accumulators, wait for the peers', reduce all of them in rank order (the same
 * bits on every GPU) and finalise (as rts_finalise_bins).  Dense bin tables only; max_bins bounds their size.
 * A rank that never arrives makes the others give up after ~2 s: RTS_ERR_STATE at the next collection, no hang. */
int rts_comm_create(rts_engine *e, uint32_t rank, uint32_t world, uint64_t max_bins);
int rts_comm_ipc_handle(rts_engine *e, void *handle_64_bytes);
int rts_comm_local_ptr(rts_engine *e, void **device_ptr);
int rts_comm_connect_ipc(rts_engine *e, const void *handles /* world x 64 bytes, rank order */);
int rts_comm_connect_ptrs(rts_engine *e, void *const *device_ptrs /* world, rank order */);
int rts_comm_allreduce_bins(rts_engine *e);
/* out = {exchanges so far, mean ns the reduce kernel waited for the slowest rank's flag, mean ns it ran}; waits for the stream. */
int rts_comm_stats(rts_engine *e, double out[3]);
int rts_comm_destroy(rts_engine *e);

/* ---- aggregation of caller-supplied received rays: the C form of rs::kernel_wrapper
 *      (aggregation.cuh:18-23, aggregation.cu:103-184).  Same array contract: accumulators
 *      pre-zeroed, path_match pre-filled (ray_tracer.cpp:1266-1271); on return rx_results[i].power,
 *      rx_results[i].doppler, delay[i], phase[i], path_match[i] are written (npath/power/doppler
 *      arrays are also written back, which the reference omits). ---- */
int rts_aggregate(rts_engine *e, rts_ray_record *rx_results, const int32_t *rx_intersects, uint32_t received,
                  uint32_t depth_total, double cspeed, double carrier, double *npath, double *power,
                  double *doppler, double *delay, double *phase, int32_t *path_match);

#ifdef __cplusplus
}
#endif
#endif /* RTS_B200_H */
