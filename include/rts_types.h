/* rts_types.h — plain-old-data carried across the rts_b200 C-ABI.
 *
 * Every struct here is POD with explicit sizes so that it can be bound from C, C++, ctypes or
 * cgo alike.  The fields are exactly the values the reference's host code reads out of the SOARS
 * `World` and hands to its OptiX programs (reference file:line given per field).  Nothing in
 * this header depends on CUDA.
 */
#ifndef RTS_TYPES_H
#define RTS_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Scene epsilons: same names and values as the reference contract (ray_tracer.h:9-10). */
#ifndef SCENE_EPS
#define SCENE_EPS 0.005f   /* minimum incident / refracted segment length, also tmin */
#endif
#ifndef SCENE_EPS_R
#define SCENE_EPS_R 0.005f /* minimum reflected segment length */
#endif

#define RTS_MAX_DEPTH 8u          /* maxRefl + maxRefr (the reference's depthTotal) upper bound  */
#define RTS_EARTH_RADIUS 6378136.0 /* ray_tracer.cu:447 */

/* Per-ray record, layout-identical to the reference's `struct PerRayData` (ray_tracer.h:13-28):
 * size 144, alignment 16, offsets asserted in rts_prd.h.  Declared here with scalar arrays so
 * that plain C and ctypes can use it; rts_prd.h provides the CUDA-vector-typed twin. */
typedef struct rts_ray_record {
    double   rayLength;        /*   0 */
    double   _pad0;            /*   8  (double2 alignment hole) */
    double   refrIndex[2];     /*  16 */
    uint32_t reflDepth;        /*  32 */
    uint32_t refrDepth;        /*  36 */
    uint32_t maxRayIndex;      /*  40 */
    uint32_t _pad1;            /*  44 */
    double   rayDirection[3];  /*  48 */
    double   firstHitPoint[3]; /*  72 */
    double   prevHitPoint[3];  /*  96 */
    double   power;            /* 120 */
    double   doppler;          /* 128 */
    int32_t  received;         /* 136 */
    uint8_t  end;              /* 140 */
    uint8_t  _pad2[3];         /* 141 */
} rts_ray_record;

/* One target's triangle mesh in world coordinates, i.e. after the reference's per-pulse
 * regenerate/rotate/translate (ray_tracer.cpp:936-1014) — the contents of dbuf_triVertices,
 * dbuf_triangles, dbuf_normals (ray_tracer.cpp:1046-1117) plus the two per-instance material
 * scalars d_targReflCoeff / d_targRefrIndex (ray_tracer.cpp:1043-1044). */
typedef struct rts_target_mesh {
    uint32_t        n_verts;
    uint32_t        n_tris;
    uint32_t        n_normals;   /* > n_verts means "per-face normals" (triangle_mesh.cu:180) */
    uint32_t        _pad;
    const double   *verts;       /* [n_verts*3]   */
    const uint32_t *tris;        /* [n_tris*3]    */
    const double   *normals;     /* [n_normals*3] */
    double          refl_coeff;  /* signed, used as-is (normal_shader.cu:298) */
    double          refr_index;
} rts_target_mesh;

/* Receiver sphere as the miss program sees it: the six dbuf_sph* / dbuf_min* / dbuf_max*
 * buffers (ray_tracer.cu:31-36), one entry per receiver. */
typedef struct rts_rx_sphere {
    double centre[3];
    double radius;
    double min_theta, max_theta;
    double min_phi, max_phi;
} rts_rx_sphere;

/* Receiver as SOARS describes it; rts_rx_sphere_from_desc() applies ray_tracer.cpp:894-918. */
typedef struct rts_rx_desc {
    double position[3];   /* Receiver::GetPosition(0)            (:902) */
    double azimuth;       /* Receiver::GetRotation(t).azimuth    (:897) */
    double elevation;     /* Receiver::GetRotation(t).elevation  (:898) */
    double radius;        /* GetRxSphere().x                     (:901) */
    double theta_span;    /* GetRxSphere().y */
    double phi_span;      /* GetRxSphere().z */
} rts_rx_desc;

/* Everything one pulse launch needs (ray_tracer.cpp:600-605, 645-648, 810-818, 881-918, 1136-1146). */
typedef struct rts_pulse {
    uint32_t nx, ny, nz;          /* launch grid. Reference: nx=ny=nz=h_numRays (:1165). An axis of
                                     extent 1 keeps the beam-start component on that axis. */
    uint32_t max_refl;            /* h_maxReflDepth (user value; device sees max_refl+1, :776)   */
    uint32_t max_refr;            /* h_maxRefrDepth; any value >0 is forced to 2 (:604-605)      */
    int32_t  interpolate_smooth;  /* rsParameters::interpolate_smooth() (:648)                  */
    double   tx_origin[3];        /* d_rayOrigin  (:881-885) */
    double   tx_dir[2];           /* d_txDir az, el (:888-889) */
    double   tx_span[3];          /* d_txSpan az span, el span, launch range (:818) */
    double   cspeed;              /* rsParameters::c() (:645) */
    double   carrier;             /* RadarSignal::GetCarrier() (:814) */
    uint32_t n_rx;
    uint32_t n_targets;           /* length of targ_vel/3; must match the committed scene */
    const rts_rx_sphere *rx;      /* [n_rx] */
    const double *targ_vel;       /* [n_targets*3]  dbuf_targ_vel (:1144-1145) */
    uint64_t ray_begin;           /* first primary-ray index of this shard            */
    uint64_t ray_count;           /* number of primary rays in this shard (0 = to end) */
    uint64_t ray_stride;          /* sample every ray_stride-th primary ray (0/1 = all) */
    /* Fused post-process (RTS_OUT_BINS) stand-ins for the SOARS callbacks of ray_tracer.cpp:1219-1247.
     * Hosts that need Target::GetRCS(angles) / antenna patterns use the two-phase path instead
     * (rts_collect_received -> host callbacks -> rts_aggregate). */
    const double *targ_rcs;       /* [n_targets] scalar RCS per target, NULL = 1 (Target::GetRCS, :1226) */
    double   gain_tx, gain_rx;    /* Gt, Gr (:1233-1235); 0 is read as 1                              */
} rts_pulse;

/* Tabulated callbacks for the fused post-process (ray_tracer.cpp:1219-1247 on the device): a function of two angles sampled
 * on a regular grid, values[i * n_el + j] = f(az0 + i * az_step, el0 + j * el_step), evaluated by bilinear interpolation
 * with the arguments clamped to the grid (no wrap-around).  n_az == 0: no table (the scalar stand-in applies). */
typedef struct rts_table2d {
    uint32_t n_az, n_el;
    double az0, az_step, el0, el_step;
    const double *values;
} rts_table2d;

/* One antenna for the fused gains (Transmitter::GetGain / Receiver::GetGain, ray_tracer.cpp:1233-1235): the gain as a
 * table over (azimuth, elevation) of the look direction MINUS the boresight, both in the simulator's spherical
 * convention SVec3(Vec3): azimuth = atan2(y, x), elevation = asin(z / length).  The boresight of a receiver is taken at
 * the ray's arrival time: bore + rate * delay (GetRotation(delay + time_t) of an antenna that turns at a constant rate).
 * position: Transmitter/Receiver::GetPosition — for a receiver NOT the capture sphere's centre (ray_tracer.cpp:1201). */
typedef struct rts_antenna {
    rts_table2d gain;
    double bore_az, bore_el, rate_az, rate_el;
    double position[3];
} rts_antenna;

/* One (receiver, target-path) group — the unit the reference's aggregation produces
 * (aggregation.cu:32-97) and from which one Response is emitted (ray_tracer.cpp:1289-1321). */
typedef struct rts_bin {
    int32_t  rx;
    int32_t  path[RTS_MAX_DEPTH];  /* dbuf_targ_intersect row; unused columns are -1 */
    int32_t  direct;               /* 1 when the members are direct rays (reflDepth==refrDepth==0) */
    double   npath;                /* d_npath_arr   (aggregation.cu:61) */
    double   sum_sqrt_power;       /* d_power_arr   (:62) */
    double   sum_delay;            /* d_delay_arr   (:63) */
    double   sum_phase;            /* d_phase_arr   (:64) */
    double   sum_doppler;          /* d_doppler_arr (:65) */
    uint64_t min_slot;             /* smallest result-slot index among the rays summed (orders like d_pathMatch, :68-69) */
    uint64_t own_min_slot;         /* smallest result-slot index among this bin's own member rays (differs for direct bins) */
    /* myKernel2 outputs (aggregation.cu:86-92) */
    double   power, delay, phase, doppler;
} rts_bin;

/* One emitted response: what ray_tracer.cpp:1301-1320 puts into InterpPoint(power, t+delay, delay,
 * doppler, phase, noise_temperature) for one unique path of one receiver. */
typedef struct rts_response {
    int32_t  rx;
    int32_t  _pad;
    uint64_t slot;                 /* representative ray (result-slot index), the unique d_pathMatch value */
    double   power, delay, doppler, phase;
} rts_response;

/* Counters returned with every pulse. */
typedef struct rts_stats {
    uint64_t primary_rays;      /* rays launched in this shard */
    uint64_t segments;          /* closest-hit queries = rtTrace calls (ray_tracer.cu:243, normal_shader.cu:268,332) */
    uint64_t hits;              /* segments that found a triangle */
    uint64_t shaded_hits;       /* hits whose closest_hit guard passed (normal_shader.cu:134) */
    uint64_t captured;          /* result slots with received >= 0 */
    uint64_t multi_captured;    /* rays captured by more than one receiver (Appendix B-Q8) */
    uint64_t edge_rays;         /* accepted hits with min(beta,gamma,1-beta-gamma) < edge epsilon */
    uint64_t refracted;         /* refracted chains spawned */
    uint64_t nodes_visited;     /* BVH nodes fetched (0 unless RTS_FLAG_COUNT_NODES) */
    uint64_t tris_tested;       /* triangle tests executed (same flag) */
    uint64_t waves;             /* wavefront launches */
    uint64_t kept_reflections;  /* first reflections answered from the hits kept from an earlier pulse (no traversal) */
    uint32_t n_bins;            /* non-empty bins */
    uint32_t primary_projected; /* 1 when the primary wave of the last batch ran by projection (RTS_RASTER=1) rather than BVH traversal */
    float    ms_update;         /* scene update + refit */
    float    ms_trace;          /* all bounce waves */
    float    ms_finalise;       /* bin finalisation */
    float    ms_total;
} rts_stats;

#ifdef __cplusplus
}
#endif
#endif /* RTS_TYPES_H */
