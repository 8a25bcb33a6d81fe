/* rts_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE, NOT A PRODUCT PATH.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load liboracle.so.  The product library
 * (rts_b200/csrc → librts_b200.so) never links, loads or calls anything in this directory.
 *
 * The oracle is a scalar C++17 restatement (OpenMP over primary rays) of the reference's
 * ray-tracing path; each function cites the reference file:line it follows.
 *
 * PARITY PIN STATUS
 *   - The reference repository holds no tests, fixtures or golden vectors, so
 *     there is nothing of the reference's own to check against directly.
 *   - Tracing stages: pinned against the reference's OWN SOURCE FILES (ray_tracer.cu,
 *     triangle_mesh.cu, normal_shader.cu) compiled unmodified for the host against an OptiX
 *     emulation shim (oracle/ref_shim → oracle/_ref/libref_rts.so) and against golden vectors
 *     generated from that build (tests/golden/, tests/golden/make_golden.py).  The shim supplies
 *     the OptiX-internal pieces the reference does not contain (reflect, refract, make_Ray,
 *     rtPotentialIntersection, traversal = exhaustive search); those are restated from the
 *     published OptiX 6.x headers from memory and remain "parity unpinned" (see DESIGN.md).
 *   - Aggregation stage: pinned on the GPU box against the reference's aggregation.cu compiled
 *     unmodified (oracle/_ref/libref_aggregation.so).
 */
#ifndef RTS_ORACLE_H
#define RTS_ORACLE_H

#include "../include/rts_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Edge-flag bits written per primary ray (any segment of any chain of that ray). */
#define ORC_EDGE_TRI     1u  /* a candidate triangle with valid t had |min(beta,gamma,1-beta-gamma)| < eps */
#define ORC_EDGE_TIE     2u  /* two accepted candidates had equal fp32 t */
#define ORC_EDGE_WINDOW  4u  /* receiver-window / pole-fold comparison within 1e-6 rad of flipping */
#define ORC_EDGE_TMIN    8u  /* a candidate's t within 4 fp32 ulp of tmin */

typedef struct orc_sizes {
    uint64_t rays;        /* nx*ny*nz                                   */
    uint64_t ray_total;   /* M*rays  (ray_tracer.cpp:608-626)           */
    uint32_t depth_total; /* D = maxRefl + maxRefr(0|2) (:655)          */
    uint32_t slots;       /* M                                          */
    uint32_t tri_cols;    /* W = maxRefl + 3 columns of the tri_path    */
    uint32_t _pad;
} orc_sizes;

int orc_sizes_for(const rts_pulse *pulse, orc_sizes *out);

/* ray_tracer.cpp:894-918 */
void orc_rx_sphere_from_desc(const rts_rx_desc *desc, rts_rx_sphere *out);

/* Full launch, reference-shaped outputs.  Any output pointer may be NULL.
 *   results        [ray_total]             dbuf_results
 *   targ_intersect [ray_total*D]           dbuf_targ_intersect (pre-filled -1)
 *   rcs_angle      [ray_total*D*2]         dbuf_rcs_angle (pre-filled -1e6)
 *   tri_path       [ray_total*W]           global triangle id per closest-hit query (-1 none)
 *   edge_flags     [rays]
 * use_bvh: 0 = exhaustive search over all triangles (definitional truth), 1 = median-split BVH.
 * Honours pulse->ray_begin / ray_count / ray_stride. */
int orc_trace(const rts_target_mesh *targets, uint32_t n_targets, const rts_pulse *pulse, int use_bvh,
              rts_ray_record *results, int32_t *targ_intersect, double *rcs_angle, int32_t *tri_path,
              uint8_t *edge_flags, rts_stats *stats);

/* The same launch with compact outputs for a shard of a large grid: n_shard = number of rays the shard selects
 * (ray_begin / ray_count / ray_stride); ray k of the shard has result slot s at index k + s*n_shard
 *   results [M*n_shard] · targ_intersect [M*n_shard*D] · rcs_angle [M*n_shard*D*2] · tri_path [M*n_shard*W] · edge_flags [n_shard] */
int orc_trace_shard(const rts_target_mesh *targets, uint32_t n_targets, const rts_pulse *pulse, int use_bvh,
                    rts_ray_record *results, int32_t *targ_intersect, double *rcs_angle, int32_t *tri_path,
                    uint8_t *edge_flags, rts_stats *stats);

/* Launch + host post-process (RCS = pulse->targ_rcs[k] or 1, Gt/Gr = pulse->gain_tx/gain_rx or 1) + binned aggregation without materialising
 * per-ray arrays.  bins sorted by (rx, path).  Returns number of bins in *n_bins (may exceed cap). */
int orc_trace_bins(const rts_target_mesh *targets, uint32_t n_targets, const rts_pulse *pulse, int use_bvh,
                   rts_bin *bins, uint32_t cap, uint32_t *n_bins, rts_stats *stats);

/* Host post-process of ray_tracer.cpp:1190-1258 with RCS = rcs_per_target[k] (NULL = 1) and
 * Gt*Gr = gain (pass 1.0): selects received slots in slot order.
 *   rx_results [cap], rx_intersects [cap*D], rx_slots [cap] (slot index of each received ray). */
int orc_postprocess(const rts_ray_record *results, const int32_t *targ_intersect, uint64_t ray_total,
                    uint32_t depth_total, double cspeed, double carrier, const double *rcs_per_target,
                    double gain, rts_ray_record *rx_results, int32_t *rx_intersects, uint64_t *rx_slots,
                    uint64_t cap, uint64_t *n_received);

/* aggregation.cu:32-97 transcribed literally (O(R^2)); arrays as rs::kernel_wrapper takes them:
 * accumulators pre-zeroed, path_match pre-filled by the caller (ray_tracer.cpp:1266-1271). */
int orc_aggregate_literal(rts_ray_record *rx_results, const int32_t *rx_intersects, uint32_t received,
                          uint32_t depth_total, double cspeed, double carrier, double *npath, double *power,
                          double *doppler, double *delay, double *phase, int32_t *path_match);

/* Same outputs by grouping on (receiver, path row): O(R log R). */
int orc_aggregate_binned(rts_ray_record *rx_results, const int32_t *rx_intersects, uint32_t received,
                         uint32_t depth_total, double cspeed, double carrier, double *npath, double *power,
                         double *doppler, double *delay, double *phase, int32_t *path_match);

/* ray_tracer.cpp:1289-1294: sort + unique of path_match. Returns count; out may be NULL. */
uint32_t orc_unique_paths(const int32_t *path_match, uint32_t received, int32_t *out);

/* ray_tracer.cpp:1289-1320 on the arrays rs::kernel_wrapper returned: one response per unique path_match
 * value u, carrying ray u's aggregated power / delay / Doppler / phase (the InterpPoint arguments) and,
 * as `slot`, rx_slots[u] (the result-slot index of that received ray).  Returns the count. */
uint32_t orc_responses(const rts_ray_record *rx_results, const double *delay, const double *phase,
                       const int32_t *path_match, const uint64_t *rx_slots, uint32_t received,
                       rts_response *out, uint32_t cap);

/* Mesh generators, ray_tracer.cpp:156-170, 226-297, 300-426, 429-504.  Two-call protocol:
 * call with NULL outputs to obtain counts, then with buffers. */
int orc_rect_mesh(float w, float h, float d, float yaw, float pitch, float roll,
                  double *verts, uint32_t *n_verts, uint32_t *tris, uint32_t *n_tris,
                  double *normals, uint32_t *n_normals);
int orc_sphere_mesh(uint32_t subdivs, float radius, float yaw, float pitch, float roll,
                    double *verts, uint32_t *n_verts, uint32_t *tris, uint32_t *n_tris,
                    double *normals, uint32_t *n_normals);
int orc_file_mesh(const char *v_file, const char *n_file, float yaw, float pitch, float roll,
                  double *verts, uint32_t *n_verts, uint32_t *tris, uint32_t *n_tris,
                  double *normals, uint32_t *n_normals);
/* vertex_rotation, ray_tracer.cpp:156-170 (float angles), in place on [n*3]. */
void orc_vertex_rotation(double *xyz, uint32_t n, float yaw, float pitch, float roll);

int orc_num_threads(void);
void orc_set_num_threads(int n);   /* OpenMP threads of the following orc_trace calls (n <= 0: unchanged) */
const char *orc_version(void);

#ifdef __cplusplus
}
#endif
#endif
