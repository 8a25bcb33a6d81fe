// ref_harness.cpp — runs the reference's OWN device programs on the CPU.
// TEST INFRASTRUCTURE.  The three reference files are #included from /root/reference (via -I),
// unmodified, each inside its own namespace (they all define the same helper names), against the
// OptiX emulation in rts_optix_shim.h.  The harness plays the role of ray_tracer.cpp's launch:
// it fills the program variables/buffers (ray_tracer.cpp:629-800, 853-918, 1020-1146) and calls
// ray_generation once per launch index of the cubic N x N x N grid (ray_tracer.cpp:1165).
//
// Single-threaded by construction (program variables are plain globals, as in OptiX's model).
#include "rts_optix_shim.h"
#include "../../include/rts_types.h"
#include <vector>

namespace ref_rg {
#include "ray_tracer.cu"
}
namespace ref_tm {
#include "triangle_mesh.cu"
}
namespace ref_ns {
#include "normal_shader.cu"
}

static_assert(sizeof(ref_rg::PerRayData) == 144 && sizeof(ref_ns::PerRayData) == 144 && sizeof(ref_tm::PerRayData) == 144,
              "PerRayData layout");
static_assert(sizeof(rts_ray_record) == 144, "rts_ray_record layout");

namespace {
struct TargetBuffers {
    std::vector<uint3> tris;
    std::vector<double3> verts, normals;
    double refl, refr;
};
struct State {
    std::vector<TargetBuffers> targets;
    // traversal state of the innermost rtTrace
    float tmin = 0, closest = 0, pending = 0;
    bool hit = false;
    unsigned cur_target = 0, cur_prim = 0;
    unsigned hit_target = 0, hit_prim = 0;
    double3 hit_normal;
    int depth = 0;
    // side channel
    int32_t *tri_path = nullptr;
    unsigned W = 0;
    uint64_t R3 = 0, rayIndex = 0;
    std::vector<unsigned> tri_offset;
    uint64_t segments = 0;
} G;

void bind_target(unsigned k)
{
    TargetBuffers &t = G.targets[k];
    ref_tm::dbuf_triangles.set(t.tris.data(), t.tris.size());
    ref_tm::dbuf_triVertices.set(t.verts.data(), t.verts.size());
    ref_tm::dbuf_normals.set(t.normals.data(), t.normals.size());
}
} // namespace

bool shim::potential_intersection(float t)
{
    if (t > G.tmin && t < G.closest) {
        G.pending = t;
        return true;
    }
    return false;
}

bool shim::report_intersection(unsigned int)
{
    G.closest = G.pending;
    G.hit = true;
    G.hit_target = G.cur_target;
    G.hit_prim = G.cur_prim;
    G.hit_normal = ref_tm::normal;
    return true;
}

void shim::trace(const optix::Ray &ray, void *payload, size_t payload_size)
{
    (void)payload_size;
    G.segments++;
    // ---- save the enclosing program's context (recursion) ----
    const optix::Ray s_ray_rg = ref_rg::ray, s_ray_tm = ref_tm::ray, s_ray_ns = ref_ns::ray;
    const ref_rg::PerRayData s_prd_rg = ref_rg::prd;
    const ref_tm::PerRayData s_prd_tm = ref_tm::prd;
    const ref_ns::PerRayData s_prd_ns = ref_ns::prd;
    const float s_hit_t = ref_ns::hit_t;
    const double3 s_normal_ns = ref_ns::normal, s_normal_tm = ref_tm::normal;
    const double s_refl = ref_ns::d_targReflCoeff, s_refr = ref_ns::d_targRefrIndex;
    const unsigned s_targ = ref_ns::d_targIndex;
    const float s_tmin = G.tmin, s_closest = G.closest, s_pending = G.pending;
    const bool s_hit = G.hit;
    const unsigned s_ht = G.hit_target, s_hp = G.hit_prim, s_ct = G.cur_target, s_cp = G.cur_prim;
    const double3 s_hn = G.hit_normal;

    // ---- the new current ray and payload ----
    unsigned char pl[144];
    memcpy(pl, payload, 144);
    ref_rg::ray = ray; ref_tm::ray = ray; ref_ns::ray = ray;
    memcpy(&ref_rg::prd, pl, 144); memcpy(&ref_tm::prd, pl, 144); memcpy(&ref_ns::prd, pl, 144);
    G.tmin = ray.tmin; G.closest = ray.tmax; G.hit = false;

    // ---- "traversal": every primitive of every target, ascending global id ----
    for (unsigned k = 0; k < G.targets.size(); k++) {
        bind_target(k);
        G.cur_target = k;
        const unsigned nprim = (unsigned)G.targets[k].tris.size();
        for (unsigned p = 0; p < nprim; p++) {
            G.cur_prim = p;
            ref_tm::intersect((int)p);
        }
    }

    unsigned char result[144];
    if (G.hit) {
        unsigned maxRayIndex;
        memcpy(&maxRayIndex, pl + 40, 4);
        if (G.tri_path && (unsigned)G.depth < G.W) {
            uint64_t slot = maxRayIndex / G.R3;
            G.tri_path[(G.rayIndex + slot * G.R3) * G.W + G.depth] = (int32_t)(G.tri_offset[G.hit_target] + G.hit_prim);
        }
        ref_ns::hit_t = G.closest;
        ref_ns::normal = G.hit_normal;
        ref_ns::d_targReflCoeff = G.targets[G.hit_target].refl;
        ref_ns::d_targRefrIndex = G.targets[G.hit_target].refr;
        ref_ns::d_targIndex = G.hit_target;
        G.depth++;
        ref_ns::closest_hit();
        G.depth--;
        memcpy(result, &ref_ns::prd, 144);
    } else {
        ref_rg::miss();
        memcpy(result, &ref_rg::prd, 144);
    }

    // ---- restore the enclosing context, then publish the payload (it may alias a restored global) ----
    ref_rg::ray = s_ray_rg; ref_tm::ray = s_ray_tm; ref_ns::ray = s_ray_ns;
    ref_rg::prd = s_prd_rg; ref_tm::prd = s_prd_tm; ref_ns::prd = s_prd_ns;
    ref_ns::hit_t = s_hit_t; ref_ns::normal = s_normal_ns; ref_tm::normal = s_normal_tm;
    ref_ns::d_targReflCoeff = s_refl; ref_ns::d_targRefrIndex = s_refr; ref_ns::d_targIndex = s_targ;
    G.tmin = s_tmin; G.closest = s_closest; G.pending = s_pending; G.hit = s_hit;
    G.hit_target = s_ht; G.hit_prim = s_hp; G.cur_target = s_ct; G.cur_prim = s_cp; G.hit_normal = s_hn;
    memcpy(payload, result, 144);
}

extern "C" const char *ref_version(void) { return "reference sources (" REF_DIR ") + OptiX emulation shim, host build"; }

// Cubic launch of the reference programs. pulse->nx must equal ny and nz (the reference has a
// single d_width).  Outputs as in orc_trace (rts_oracle.h); any may be NULL except results.
extern "C" int ref_trace(const rts_target_mesh *targets, uint32_t n_targets, const rts_pulse *p,
                         rts_ray_record *results, int32_t *targ_intersect, double *rcs_angle, int32_t *tri_path,
                         uint64_t *segments)
{
    if (!p || !results) return -1;
    if (p->nx != p->ny || p->ny != p->nz) return -2;
    const unsigned N = p->nx;
    const unsigned maxRefl = p->max_refl;
    const unsigned maxRefr = p->max_refr > 0 ? 2 : 0;                 // ray_tracer.cpp:604-605
    unsigned M = 1;
    if (maxRefr == 2) M += (maxRefl + 1) + 1;                          // ray_tracer.cpp:608-613
    const uint64_t R3 = (uint64_t)N * N * N;
    const uint64_t rayTotal = M * R3;                                  // ray_tracer.cpp:626
    const unsigned D = maxRefr + maxRefl;                              // ray_tracer.cpp:655
    const unsigned W = maxRefl + 3;

    // scene
    G.targets.assign(n_targets, TargetBuffers());
    G.tri_offset.assign(n_targets, 0);
    unsigned off = 0;
    for (uint32_t k = 0; k < n_targets; k++) {
        TargetBuffers &t = G.targets[k];
        G.tri_offset[k] = off;
        off += targets[k].n_tris;
        t.tris.resize(targets[k].n_tris);
        for (uint32_t i = 0; i < targets[k].n_tris; i++)
            t.tris[i] = make_uint3(targets[k].tris[3 * i], targets[k].tris[3 * i + 1], targets[k].tris[3 * i + 2]);
        t.verts.resize(targets[k].n_verts);
        for (uint32_t i = 0; i < targets[k].n_verts; i++)
            t.verts[i] = make_double3(targets[k].verts[3 * i], targets[k].verts[3 * i + 1], targets[k].verts[3 * i + 2]);
        t.normals.resize(targets[k].n_normals);
        for (uint32_t i = 0; i < targets[k].n_normals; i++)
            t.normals[i] = make_double3(targets[k].normals[3 * i], targets[k].normals[3 * i + 1], targets[k].normals[3 * i + 2]);
        t.refl = targets[k].refl_coeff;
        t.refr = targets[k].refr_index;
    }

    // host-side buffer defaults (ray_tracer.cpp:854-868)
    std::vector<int32_t> ti_local;
    std::vector<double2> rcs_local((size_t)rayTotal * (D ? D : 1));
    if (!targ_intersect) { ti_local.resize((size_t)rayTotal * (D ? D : 1)); targ_intersect = ti_local.data(); }
    for (uint64_t i = 0; i < rayTotal * D; i++) targ_intersect[i] = -1;
    for (uint64_t i = 0; i < rayTotal * D; i++) { rcs_local[i].x = -1000000; rcs_local[i].y = -1000000; }
    if (tri_path) for (uint64_t i = 0; i < rayTotal * W; i++) tri_path[i] = -1;
    memset(results, 0, sizeof(rts_ray_record) * rayTotal);

    // receivers (ray_tracer.cu:31-36)
    std::vector<double3> sc(p->n_rx);
    std::vector<double> sr(p->n_rx), mint(p->n_rx), maxt(p->n_rx), minp(p->n_rx), maxp(p->n_rx);
    for (uint32_t j = 0; j < p->n_rx; j++) {
        sc[j] = make_double3(p->rx[j].centre[0], p->rx[j].centre[1], p->rx[j].centre[2]);
        sr[j] = p->rx[j].radius;
        mint[j] = p->rx[j].min_theta; maxt[j] = p->rx[j].max_theta;
        minp[j] = p->rx[j].min_phi; maxp[j] = p->rx[j].max_phi;
    }
    ref_rg::dbuf_sphCentre.set(sc.data(), sc.size());
    ref_rg::dbuf_sphRadius.set(sr.data(), sr.size());
    ref_rg::dbuf_minTheta.set(mint.data(), mint.size());
    ref_rg::dbuf_maxTheta.set(maxt.data(), maxt.size());
    ref_rg::dbuf_minPhi.set(minp.data(), minp.size());
    ref_rg::dbuf_maxPhi.set(maxp.data(), maxp.size());

    // target velocities (ray_tracer.cpp:1136-1146)
    std::vector<double3> vel(n_targets);
    for (uint32_t k = 0; k < n_targets; k++)
        vel[k] = p->targ_vel ? make_double3(p->targ_vel[3 * k], p->targ_vel[3 * k + 1], p->targ_vel[3 * k + 2])
                             : make_double3(0, 0, 0);
    ref_ns::dbuf_targ_vel.set(vel.data(), vel.size());

    // outputs
    ref_rg::dbuf_results.set((ref_rg::PerRayData *)results, rayTotal);
    ref_ns::dbuf_results.set((ref_ns::PerRayData *)results, rayTotal);
    ref_ns::dbuf_targ_intersect.set(targ_intersect, D, rayTotal);
    ref_ns::dbuf_rcs_angle.set(rcs_local.data(), D, rayTotal);

    // context variables (ray_tracer.cpp:774-799, 818, 881-889)
    ref_rg::d_width = N; ref_ns::d_width = N;
    ref_rg::d_rxsize = p->n_rx;
    ref_rg::d_maxRayTotal = (unsigned)rayTotal;
    ref_ns::d_maxReflDepth = maxRefl + 1;                               // ray_tracer.cpp:776
    ref_ns::d_maxRefrDepth = maxRefr;
    ref_tm::d_interpolate_smooth = p->interpolate_smooth != 0;
    const double3 origin = make_double3(p->tx_origin[0], p->tx_origin[1], p->tx_origin[2]);
    ref_rg::d_rayOrigin = origin; ref_ns::d_rayOrigin = origin;
    ref_rg::d_txSpan = make_double3(p->tx_span[0], p->tx_span[1], p->tx_span[2]);
    ref_rg::d_txDir = make_double2(p->tx_dir[0], p->tx_dir[1]);

    G.tri_path = tri_path; G.W = W; G.R3 = R3; G.depth = 0; G.segments = 0;

    for (unsigned z = 0; z < N; z++)
        for (unsigned y = 0; y < N; y++)
            for (unsigned x = 0; x < N; x++) {
                ref_rg::launchIndex = make_uint3(x, y, z);
                ref_ns::launchIndex = make_uint3(x, y, z);
                G.rayIndex = (uint64_t)z * N * N + (uint64_t)y * N + x;
                G.depth = 0;
                ref_rg::ray_generation();
            }

    if (rcs_angle)
        for (uint64_t i = 0; i < rayTotal * D; i++) { rcs_angle[2 * i] = rcs_local[i].x; rcs_angle[2 * i + 1] = rcs_local[i].y; }
    if (segments) *segments = G.segments;
    return 0;
}

// The reference's bounding-box program for one target's triangles (triangle_mesh.cu:204-233).
extern "C" int ref_bounds(const rts_target_mesh *target, float *out6 /* [n_tris*6] */)
{
    G.targets.assign(1, TargetBuffers());
    TargetBuffers &t = G.targets[0];
    t.tris.resize(target->n_tris);
    for (uint32_t i = 0; i < target->n_tris; i++)
        t.tris[i] = make_uint3(target->tris[3 * i], target->tris[3 * i + 1], target->tris[3 * i + 2]);
    t.verts.resize(target->n_verts);
    for (uint32_t i = 0; i < target->n_verts; i++)
        t.verts[i] = make_double3(target->verts[3 * i], target->verts[3 * i + 1], target->verts[3 * i + 2]);
    bind_target(0);
    for (uint32_t i = 0; i < target->n_tris; i++) ref_tm::bound((int)i, out6 + 6 * (size_t)i);
    return 0;
}
