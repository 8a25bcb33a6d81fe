// oracle_trace.cpp — scalar CPU restatement of the reference's tracing programs.
// TEST INFRASTRUCTURE (see rts_oracle.h).  Must be compiled with -ffp-contract=off.
//
// Follows, statement by statement:
//   ray generation            /root/reference/ray_tracer.cu:144-255
//   miss / receiver capture   /root/reference/ray_tracer.cu:260-478
//   triangle test + normal    /root/reference/triangle_mesh.cu:121-200
//   closest hit               /root/reference/normal_shader.cu:128-340
//   sizing / buffer defaults  /root/reference/ray_tracer.cpp:600-626,655,776-778,854-868
// OptiX-internal functions the reference calls but does not contain (reflect, refract,
// normalize, make_Ray's tmax, rtPotentialIntersection) are isolated in the `optix_recalled`
// namespace below: restated from the published OptiX 6.x optixu_math_namespace.h from memory
// ("parity unpinned" for exactly those functions).
#include "oracle_common.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <map>
#include <omp.h>

namespace orc {

// ------------------------------------------------------------------------------------------
namespace optix_recalled {
static const float RT_DEFAULT_MAX = 1.e27f;
static inline float dotf(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline F3 normalizef(F3 v)
{
    float invLen = 1.0f / sqrtf(dotf(v, v));
    return F3{v.x * invLen, v.y * invLen, v.z * invLen};
}
// reflect(i, n) = i - 2.0f * n * dot(n, i)
static inline F3 reflect(F3 i, F3 n)
{
    float d = dotf(n, i);
    return F3{i.x - (2.0f * n.x) * d, i.y - (2.0f * n.y) * d, i.z - (2.0f * n.z) * d};
}
// refract(r, i, n, ior): Snell with the side chosen from sign(dot(i, n)); false on TIR.
static inline bool refract(F3 &r, F3 i, F3 n, float ior)
{
    F3 nn = n;
    float negNdotV = dotf(i, nn);
    float eta;
    if (negNdotV > 0.0f) {
        eta = ior;
        nn = F3{-n.x, -n.y, -n.z};
        negNdotV = -negNdotV;
    } else {
        eta = 1.f / ior;
    }
    const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
    if (k < 0.0f) {
        r = F3{0.f, 0.f, 0.f};
        return false;
    }
    const float s = eta * negNdotV + sqrtf(k);
    r = normalizef(F3{eta * i.x - s * nn.x, eta * i.y - s * nn.y, eta * i.z - s * nn.z});
    return true;
}
} // namespace optix_recalled

// ------------------------------------------------------------------------------------------
struct RayF { F3 origin; F3 direction; float tmin; float tmax; };

struct Launch {
    const Scene *scene;
    const rts_pulse *p;
    uint32_t nx, ny, nz;
    uint64_t R3;
    uint32_t dMax;   // d_maxReflDepth = max_refl + 1 (ray_tracer.cpp:776)
    uint32_t rMax;   // d_maxRefrDepth in {0,2}      (ray_tracer.cpp:604-605)
    uint32_t D;      // max_refl + rMax              (ray_tracer.cpp:655)
    uint32_t M;      // result slots per primary ray (ray_tracer.cpp:608-613)
    uint32_t W;      // tri_path columns
    D3 origin;
    bool interpolate;
    // launch-invariant pieces of ray generation, hoisted (same expressions, same order)
    D3 beamStart, beamEnd;
    double slope[3];
    double Rot[3][3], Rot1[3][3];
    D3 boresight;
};

// Per-primary-ray output window: slot k of this ray lives at index k*stride from the bases.
struct RayOut {
    rts_ray_record *res;
    int32_t *ti;      // rows of D
    double *rcs;      // rows of 2*D
    int32_t *tp;      // rows of W
    uint64_t res_stride;  // record index distance between consecutive slots
    uint64_t row_stride;  // row index distance between consecutive slots
    uint8_t edge;
    // counters
    uint64_t segments, hits, shaded, refracted, multi;
};

static inline D3 sph_to_cart(double azi, double ele) // ray_tracer.cu:132-139
{
    D3 c;
    c.x = cos(azi) * cos(ele);
    c.y = sin(azi) * cos(ele);
    c.z = sin(ele);
    return c;
}

static void setup_launch(Launch &L, const Scene &scene, const rts_pulse *p)
{
    L.scene = &scene;
    L.p = p;
    L.nx = p->nx; L.ny = p->ny; L.nz = p->nz;
    L.R3 = (uint64_t)p->nx * p->ny * p->nz;
    L.dMax = p->max_refl + 1;
    L.rMax = p->max_refr > 0 ? 2u : 0u;
    L.D = p->max_refl + L.rMax;
    L.M = L.rMax == 2 ? 1 + (p->max_refl + 1) + 1 : 1;
    L.W = p->max_refl + 3;
    L.origin = d3(p->tx_origin[0], p->tx_origin[1], p->tx_origin[2]);
    L.interpolate = p->interpolate_smooth != 0;
    const double az = p->tx_dir[0], el = p->tx_dir[1];
    // ray_tracer.cu:155-156
    L.beamStart = sph_to_cart(-p->tx_span[0] / 2, -p->tx_span[1] / 2);
    L.beamEnd = sph_to_cart(p->tx_span[0] / 2, p->tx_span[1] / 2);
    // ray_tracer.cu:167-169 (per-axis extent generalises d_width; extent 1 → no slope term)
    L.slope[0] = L.nx > 1 ? (((L.beamEnd.x * (1 + p->tx_span[2])) - L.beamStart.x) / (L.nx - 1)) : 0.0;
    L.slope[1] = L.ny > 1 ? ((L.beamEnd.y - L.beamStart.y) / (L.ny - 1)) : 0.0;
    L.slope[2] = L.nz > 1 ? ((L.beamEnd.z - L.beamStart.z) / (L.nz - 1)) : 0.0;
    // ray_tracer.cu:173-175
    double Rot[3][3] = {{cos(az), -sin(az), 0}, {sin(az), cos(az), 0}, {0, 0, 1}};
    memcpy(L.Rot, Rot, sizeof(Rot));
    // ray_tracer.cu:186-190
    D3 rotated = d3(0, 0, 0);
    rotated.x += Rot[0][1];
    rotated.y += Rot[1][1];
    rotated.z += Rot[2][1];
    D3 o = normalised(rotated);
    // ray_tracer.cu:194-196 (sin signs as in the reference)
    double c = cos(el), s = sin(el);
    double Rot1[3][3] = {
        {c + o.x * o.x * (1 - c), o.x * o.y * (1 - c) + o.z * s, o.x * o.z * (1 - c) - o.y * s},
        {o.y * o.x * (1 - c) - o.z * s, c + o.y * o.y * (1 - c), o.y * o.z * (1 - c) + o.x * s},
        {o.z * o.x * (1 - c) + o.y * s, o.z * o.y * (1 - c) - o.x * s, c + o.z * o.z * (1 - c)}};
    memcpy(L.Rot1, Rot1, sizeof(Rot1));
    L.boresight = sph_to_cart(az, el); // ray_tracer.cu:161
}

// ray_tracer.cu:158-204
static D3 primary_direction(const Launch &L, uint32_t ix, uint32_t iy, uint32_t iz)
{
    if (L.nx == 1 && L.ny == 1 && L.nz == 1) return L.boresight;
    D3 d;
    d.x = L.nx > 1 ? L.beamStart.x + L.slope[0] * (ix) : L.beamStart.x;
    d.y = L.ny > 1 ? L.beamStart.y + L.slope[1] * (iy) : L.beamStart.y;
    d.z = L.nz > 1 ? L.beamStart.z + L.slope[2] * (iz) : L.beamStart.z;
    d = normalised(d);
    D3 r = d3(0, 0, 0);
    r.x += L.Rot[0][0] * d.x + L.Rot[0][1] * d.y + L.Rot[0][2] * d.z;
    r.y += L.Rot[1][0] * d.x + L.Rot[1][1] * d.y + L.Rot[1][2] * d.z;
    r.z += L.Rot[2][0] * d.x + L.Rot[2][1] * d.y + L.Rot[2][2] * d.z;
    d = normalised(r);
    r = d3(0, 0, 0);
    r.x += L.Rot1[0][0] * d.x + L.Rot1[0][1] * d.y + L.Rot1[0][2] * d.z;
    r.y += L.Rot1[1][0] * d.x + L.Rot1[1][1] * d.y + L.Rot1[1][2] * d.z;
    r.z += L.Rot1[2][0] * d.x + L.Rot1[2][1] * d.y + L.Rot1[2][2] * d.z;
    return r; // not re-normalised (ray_tracer.cu:203)
}

// ------------------------------------------------------------------------------------------
// Closest-hit selection: triangle_mesh.cu:121-137 per candidate + rtPotentialIntersection.
struct Hit {
    int64_t tri;   // global triangle id, -1 = none
    uint32_t mesh, prim;
    float t;       // hit_t as the shader sees it (fp32)
    double beta, gamma;
    D3 n;          // unnormalised geometric normal of the winner
};

static const double EDGE_EPS = 1e-9;

static inline void test_triangle(const Mesh &m, uint32_t prim, const D3 &o, const D3 &d, const RayF &ray,
                                 Hit &best, uint8_t &edge)
{
    const uint32_t *t3 = m.tris + 3 * (size_t)prim;
    const D3 p0 = d3(m.verts[3 * (size_t)t3[0]], m.verts[3 * (size_t)t3[0] + 1], m.verts[3 * (size_t)t3[0] + 2]);
    const D3 p1 = d3(m.verts[3 * (size_t)t3[1]], m.verts[3 * (size_t)t3[1] + 1], m.verts[3 * (size_t)t3[1] + 2]);
    const D3 p2 = d3(m.verts[3 * (size_t)t3[2]], m.verts[3 * (size_t)t3[2] + 1], m.verts[3 * (size_t)t3[2] + 2]);
    const D3 e0 = sub(p1, p0);
    const D3 e1 = sub(p0, p2);
    const D3 n = cross(e1, e0);
    const D3 e2 = scale((1 / dot(n, d)), sub(p0, o));
    const D3 i = cross(d, e2);
    const double beta = dot(i, e1);
    const double gamma = dot(i, e0);
    const double t = dot(n, e2);
    const bool in_t = (t < ray.tmax) & (t > ray.tmin);
    if (in_t) {
        double mb = std::min(std::min(beta, gamma), 1 - beta - gamma);
        if (std::fabs(mb) < EDGE_EPS) edge |= ORC_EDGE_TRI;
        if (std::fabs((double)(float)t - (double)ray.tmin) <= 4 * 4.7e-10) edge |= ORC_EDGE_TMIN;
    }
    if (in_t & (beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1)) {
        // rtPotentialIntersection(t): t narrowed to float, strictly inside (tmin, current closest)
        const float tf = (float)t;
        const int64_t gid = (int64_t)m.tri_offset + prim;
        if (tf > ray.tmin) {
            if (best.tri >= 0 && tf == best.t) edge |= ORC_EDGE_TIE;
            // tie rule: smaller fp32 t, then lower global triangle id
            if (best.tri < 0 || tf < best.t || (tf == best.t && gid < best.tri)) {
                best.tri = gid;
                best.mesh = 0; // filled by caller
                best.prim = prim;
                best.t = tf;
                best.beta = beta;
                best.gamma = gamma;
                best.n = n;
            }
        }
    }
}

static Hit closest_brute(const Scene &s, const RayF &ray, const D3 &o, const D3 &d, uint8_t &edge)
{
    Hit best;
    best.tri = -1; best.t = ray.tmax; best.mesh = 0; best.prim = 0; best.beta = best.gamma = 0; best.n = d3(0, 0, 0);
    for (uint32_t mi = 0; mi < s.meshes.size(); mi++) {
        const Mesh &m = s.meshes[mi];
        for (uint32_t prim = 0; prim < m.n_tris; prim++) test_triangle(m, prim, o, d, ray, best, edge);
    }
    if (best.tri >= 0) best.mesh = s.tri_mesh[(size_t)best.tri];
    return best;
}

static Hit closest_bvh(const Scene &s, const RayF &ray, const D3 &o, const D3 &d, uint8_t &edge)
{
    Hit best;
    best.tri = -1; best.t = ray.tmax; best.mesh = 0; best.prim = 0; best.beta = best.gamma = 0; best.n = d3(0, 0, 0);
    const double inv[3] = {1.0 / d.x, 1.0 / d.y, 1.0 / d.z};
    const double oo[3] = {o.x, o.y, o.z};
    int32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const BvhNode &nd = s.nodes[stack[--sp]];
        // slab test in fp64, padded generously (conservative: never prunes a box the exact
        // triangle test could accept; NaN from 0*inf is ignored by fmin/fmax)
        // The box is first grown by delta = 1e-12 * (|box| + |o|) per axis: the fp64 triangle test accepts hits a few
        // ulps outside the exact triangle (a ray running along the shared edge of two triangles, d.y ~ 1e-17), and
        // a box that ends exactly on that edge must not prune them — exhaustive search is the truth this mode has
        // to reproduce (tests/test_oracle_bvh.py).
        double tn = -INFINITY, tf = INFINITY;
        for (int a = 0; a < 3; a++) {
            const double delta = 1e-12 * (fmax(fabs(nd.lo[a]), fabs(nd.hi[a])) + fabs(oo[a])) + 1e-300;
            double t0 = ((nd.lo[a] - delta) - oo[a]) * inv[a];
            double t1 = ((nd.hi[a] + delta) - oo[a]) * inv[a];
            tn = fmax(tn, fmin(fmin(t0, t1), INFINITY));
            tf = fmin(tf, fmax(fmax(t0, t1), -INFINITY));
        }
        const double pad = 1e-6;
        tn = tn - fabs(tn) * pad - 1e-9;
        tf = tf + fabs(tf) * pad + 1e-9;
        const double tbest = best.tri >= 0 ? (double)best.t * (1 + pad) + 1e-9 : INFINITY;
        if (!(tn <= tf) || tf < 0 || tn > tbest) continue;
        if (nd.count > 0) {
            for (int k = 0; k < nd.count; k++) {
                uint32_t gid = s.order[nd.start + k];
                const Mesh &m = s.meshes[s.tri_mesh[gid]];
                test_triangle(m, gid - m.tri_offset, o, d, ray, best, edge);
            }
        } else {
            stack[sp++] = nd.left;
            stack[sp++] = nd.right;
        }
    }
    if (best.tri >= 0) best.mesh = s.tri_mesh[(size_t)best.tri];
    return best;
}

// Normal attribute, triangle_mesh.cu:170-194
static D3 hit_normal(const Launch &L, const Hit &h)
{
    const Mesh &m = L.scene->meshes[h.mesh];
    if (L.interpolate) {
        D3 normal;
        if (m.n_normals > m.n_verts) {
            normal = d3(m.normals[3 * (size_t)h.prim], m.normals[3 * (size_t)h.prim + 1], m.normals[3 * (size_t)h.prim + 2]);
        } else {
            const uint32_t *t3 = m.tris + 3 * (size_t)h.prim;
            auto nv = [&](uint32_t v) {
                if (v >= m.n_normals) return d3(0, 0, 0);
                return d3(m.normals[3 * (size_t)v], m.normals[3 * (size_t)v + 1], m.normals[3 * (size_t)v + 2]);
            };
            const D3 n0 = nv(t3[0]), n1 = nv(t3[1]), n2 = nv(t3[2]);
            const double beta = h.beta, gamma = h.gamma;
            normal = d3(n1.x * beta + n2.x * gamma + n0.x * (1.0f - beta - gamma),
                        n1.y * beta + n2.y * gamma + n0.y * (1.0f - beta - gamma),
                        n1.z * beta + n2.z * gamma + n0.z * (1.0f - beta - gamma));
        }
        return normalised(normal);
    }
    return normalised(h.n);
}

// ------------------------------------------------------------------------------------------
static void trace(const Launch &L, const RayF &ray, rts_ray_record &prd, uint64_t rayIndex, uint32_t col, RayOut &out,
                  bool use_bvh);

static inline D3 rd3(const double *v) { return d3(v[0], v[1], v[2]); }
static inline void wd3(double *v, D3 a) { v[0] = a.x; v[1] = a.y; v[2] = a.z; }

static inline void cart_to_sph(D3 in, double &azi, double &ele) // normal_shader.cu:118-124
{
    azi = atan2(in.y, in.x);
    ele = atan2(in.z, sqrt(in.x * in.x + in.y * in.y));
}

static inline void write_back(rts_ray_record &dst, const rts_ray_record &src) // ray_tracer.cu:246-253, normal_shader.cu:272-279
{
    dst.reflDepth = src.reflDepth;
    dst.refrDepth = src.refrDepth;
    dst.rayLength = src.rayLength;
    memcpy(dst.firstHitPoint, src.firstHitPoint, sizeof(double) * 3);
    memcpy(dst.prevHitPoint, src.prevHitPoint, sizeof(double) * 3);
    dst.power = src.power;
    dst.doppler = src.doppler;
    dst.received = src.received;
}

// normal_shader.cu:128-340
static void closest_hit(const Launch &L, const RayF &ray, rts_ray_record &prd, const Hit &h, uint64_t rayIndex,
                        uint32_t col, RayOut &out, bool use_bvh)
{
    const uint32_t d_maxReflDepth = L.dMax, d_maxRefrDepth = L.rMax;
    if ((prd.end == false) && ((prd.refrDepth < d_maxRefrDepth) || (prd.reflDepth < (d_maxReflDepth - 1)))) {
        out.shaded++;
        const Mesh &mesh = L.scene->meshes[h.mesh];
        const uint32_t d_targIndex = h.mesh;
        const double d_targReflCoeff = mesh.refl_coeff, d_targRefrIndex = mesh.refr_index;
        const float hit_t = h.t;
        const D3 normal = hit_normal(L, h);
        const uint64_t R3 = L.R3;
        (void)rayIndex;

        // :140-146  path row
        if (prd.refrDepth != 1) {
            uint64_t slot = prd.maxRayIndex / R3; // rayIndex + maxRayIndex  → slot number
            uint32_t x = prd.reflDepth + prd.refrDepth;
            if (x < (d_maxRefrDepth + d_maxReflDepth - 1))
                if (out.ti) out.ti[(slot * out.row_stride) * L.D + x] = (int)(d_targIndex);
        }

        // :148-152
        D3 prev = rd3(prd.prevHitPoint), dir = rd3(prd.rayDirection);
        D3 hitPoint;
        hitPoint.x = prev.x + (double)hit_t * dir.x;
        hitPoint.y = prev.y + (double)hit_t * dir.y;
        hitPoint.z = prev.z + (double)hit_t * dir.z;
        prd.rayLength += hit_t;

        // :158-173
        if ((prd.reflDepth == 0) && (prd.refrDepth == 0)) {
            wd3(prd.firstHitPoint, hitPoint);
            D3 TxRange = sub(rd3(prd.firstHitPoint), L.origin);
            if (length(TxRange) >= SCENE_EPS)
                prd.power = 1 / ((magsq(TxRange)) * 4 * M_PI);
            else
                prd.end = true;
        } else {
            D3 TargRange = sub(hitPoint, rd3(prd.prevHitPoint));
            if (length(TargRange) >= SCENE_EPS_R)
                prd.power *= 1 / ((magsq(TargRange)) * 4 * M_PI);
            else
                prd.end = true;
        }

        // :176
        wd3(prd.prevHitPoint, hitPoint);

        // :183-188
        F3 hitPoint_f3 = F3{(float)hitPoint.x, (float)hitPoint.y, (float)hitPoint.z};
        F3 new_direction;
        const double *vt = L.p->targ_vel + 3 * (size_t)d_targIndex;
        D3 V_targ = d3(vt[0], vt[1], vt[2]);
        D3 k1, k0;

        // :191-194
        rts_ray_record prd_refr = prd;
        prd_refr.refrIndex[0] = prd_refr.refrIndex[1];

        // :198
        if ((fabs(d_targReflCoeff) != 1.00000f) && (prd_refr.refrDepth < d_maxRefrDepth) && (prd_refr.reflDepth == 0)) {
            // :201-206
            if (prd_refr.refrIndex[0] == 1)
                prd_refr.refrIndex[1] = d_targRefrIndex;
            else
                prd_refr.refrIndex[1] = 1;
            // :209
            float refr_index_ratio = (float)(prd_refr.refrIndex[1] / prd_refr.refrIndex[0]);
            // :212
            if (optix_recalled::refract(new_direction, ray.direction, normalise_float3(normal.x, normal.y, normal.z),
                                        refr_index_ratio)) {
                out.refracted++;
                // :214-215 (slot arithmetic in 64 bit; the record field stays 32 bit)
                uint64_t currentRayIndex = (uint64_t)prd_refr.maxRayIndex + R3;
                prd_refr.maxRayIndex = (uint32_t)currentRayIndex;
                const uint64_t curSlot = currentRayIndex / R3;

                // :221-239 pre-fill
                if ((prd_refr.refrDepth == 0) && (currentRayIndex == R3)) {
                    if (out.ti) {
                        for (uint32_t i = 0; i < (d_maxReflDepth + d_maxRefrDepth - 1); i++)
                            out.ti[(1 * out.row_stride) * L.D + i] = (int)(d_targIndex);
                        for (uint32_t j = 0; j < d_maxReflDepth; j++)
                            for (uint32_t i = 0; i < (j + 2); i++)
                                if (i < L.D) out.ti[((uint64_t)(j + 2) * out.row_stride) * L.D + i] = (int)(d_targIndex);
                    }
                }

                // :242
                RayF refr_ray{hitPoint_f3, new_direction, SCENE_EPS, optix_recalled::RT_DEFAULT_MAX};

                // :245-247
                if ((prd_refr.reflDepth + 1) < d_maxReflDepth) prd_refr.power *= (1 - fabs(d_targReflCoeff));
                prd_refr.refrDepth++;

                // :251-256
                k0 = normalised(rd3(prd_refr.rayDirection));
                wd3(prd_refr.rayDirection, d3(new_direction.x, new_direction.y, new_direction.z));
                k1 = normalised(rd3(prd_refr.rayDirection));
                prd_refr.doppler += dot(V_targ, sub(k1, k0));

                // :259-265
                {
                    uint32_t x = prd_refr.reflDepth + (prd_refr.refrDepth - 1);
                    double a0, e0, a1, e1;
                    cart_to_sph(k0, a0, e0);
                    cart_to_sph(d3(-k1.x, -k1.y, -k1.z), a1, e1);
                    if (out.rcs && x < L.D) {
                        out.rcs[((curSlot * out.row_stride) * L.D + x) * 2 + 0] = a0 + a1;
                        out.rcs[((curSlot * out.row_stride) * L.D + x) * 2 + 1] = e0 + e1;
                    }
                }

                // :268
                trace(L, refr_ray, prd_refr, rayIndex, col + 1, out, use_bvh);

                // :272-279
                if (out.res) write_back(out.res[curSlot * out.res_stride], prd_refr);
            }
        }

        // :286-290
        prd.reflDepth++;
        prd.refrIndex[1] = prd_refr.refrIndex[0];
        prd.refrIndex[0] = prd_refr.refrIndex[0];

        // :293-333
        if (prd.reflDepth < d_maxReflDepth) {
            new_direction = optix_recalled::reflect(ray.direction, normalise_float3(normal.x, normal.y, normal.z));
            RayF refl_ray{hitPoint_f3, new_direction, SCENE_EPS_R, optix_recalled::RT_DEFAULT_MAX};
            prd.power *= d_targReflCoeff;

            k0 = normalised(rd3(prd.rayDirection));
            wd3(prd.rayDirection, d3(new_direction.x, new_direction.y, new_direction.z));
            k1 = normalised(rd3(prd.rayDirection));
            prd.doppler += dot(V_targ, sub(k1, k0));

            {
                uint64_t slot = prd.maxRayIndex / R3;
                uint32_t x = (prd.reflDepth - 1) + prd.refrDepth;
                double a0, e0, a1, e1;
                cart_to_sph(k0, a0, e0);
                cart_to_sph(d3(-k1.x, -k1.y, -k1.z), a1, e1);
                if (out.rcs && x < L.D) {
                    out.rcs[((slot * out.row_stride) * L.D + x) * 2 + 0] = a0 + a1;
                    out.rcs[((slot * out.row_stride) * L.D + x) * 2 + 1] = e0 + e1;
                }
            }
            trace(L, refl_ray, prd, rayIndex, col + 1, out, use_bvh);
        }

        // :336-338
        if ((prd.reflDepth + 1 >= d_maxReflDepth) && (prd.refrDepth >= d_maxRefrDepth)) prd.end = true;
    }
}

// ray_tracer.cu:53-69
static inline void normalise_angle(double &angle)
{
    while (angle < -M_PI) angle += 2 * M_PI;
    while (angle > M_PI) angle -= 2 * M_PI;
}
static inline bool angle_in_range(double testAngle, double a, double b, uint8_t &edge)
{
    a -= testAngle;
    b -= testAngle;
    normalise_angle(a);
    normalise_angle(b);
    const double delta = 1e-6;
    if (fabs(a) < delta || fabs(b) < delta || fabs(fabs(a) - M_PI) < delta || fabs(fabs(b) - M_PI) < delta ||
        fabs(fabs(a - b) - M_PI) < delta)
        edge |= ORC_EDGE_WINDOW;
    if (a * b >= 0) return false;
    return fabs(a - b) < M_PI;
}

// ray_tracer.cu:260-478
static void miss(const Launch &L, rts_ray_record &prd, RayOut &out)
{
    const rts_pulse *p = L.p;
    const D3 o = rd3(prd.prevHitPoint), dir = rd3(prd.rayDirection);
    if (prd.end == false) {
        double A, B, C, discriminant;
        double t[2] = {0, 0};
        uint32_t captures = 0;
        for (unsigned int Rx_i = 0; Rx_i < p->n_rx; Rx_i++) {
            const rts_rx_sphere &rx = p->rx[Rx_i];
            const D3 c = d3(rx.centre[0], rx.centre[1], rx.centre[2]);
            // prd fields may have been modified by an earlier receiver in this loop (Q8)
            const D3 po = rd3(prd.prevHitPoint);
            (void)o;
            A = (dir.x) * (dir.x) + (dir.y) * (dir.y) + (dir.z) * (dir.z);
            B = 2 * (((po.x - c.x) * dir.x) + ((po.y - c.y) * dir.y) + ((po.z - c.z) * dir.z));
            C = po.x * po.x + po.y * po.y + po.z * po.z + (c.x * c.x) + (c.y * c.y) + (c.z * c.z) -
                2 * ((c.x * po.x) + (c.y * po.y) + (c.z * po.z)) - rx.radius * rx.radius;
            discriminant = B * B - 4 * A * C;
            if (discriminant > 0.f) {
                discriminant = sqrt(discriminant);
                t[0] = (-B - discriminant) / (2 * A);
                t[1] = (-B + discriminant) / (2 * A);
                unsigned int received_root = 2;
                for (int i = 0; i < 2; i++) {
                    if ((t[i] >= 0) && ((prd.rayLength + t[i]) > SCENE_EPS) && ((prd.rayLength + t[i]) > SCENE_EPS_R)) {
                        D3 endPoint;
                        endPoint.x = po.x + t[i] * dir.x;
                        endPoint.y = po.y + t[i] * dir.y;
                        endPoint.z = po.z + t[i] * dir.z;
                        double theta = atan2f((endPoint.y - c.y), (endPoint.x - c.x));
                        double phi = atan2f(endPoint.z - c.z, sqrt(((endPoint.y - c.y) * (endPoint.y - c.y)) +
                                                                   ((endPoint.x - c.x) * (endPoint.x - c.x))));
                        if (fabs(fabs(phi) - M_PI / 2) < 1e-6) out.edge |= ORC_EDGE_WINDOW;
                        if ((phi < -M_PI / 2)) {
                            theta += M_PI;
                            phi = -M_PI - phi;
                        }
                        if ((phi > M_PI / 2)) {
                            theta += M_PI;
                            phi = M_PI - phi;
                        }
                        double d_maxTheta1 = rx.max_theta;
                        double d_minTheta1 = rx.min_theta;
                        double d_maxTheta2 = d_maxTheta1;
                        double d_minTheta2 = d_minTheta1;
                        double d_maxPhi1 = rx.max_phi;
                        double d_minPhi1 = rx.min_phi;
                        double d_maxPhi2 = d_maxPhi1;
                        double d_minPhi2 = d_minPhi1;
                        if ((d_minPhi1 < -M_PI / 2)) {
                            d_maxTheta2 += M_PI;
                            d_minTheta2 += M_PI;
                            d_maxPhi2 = -M_PI - d_minPhi1;
                            d_minPhi2 = -M_PI / 2;
                            d_minPhi1 = -M_PI / 2;
                        }
                        if ((d_maxPhi1 > M_PI / 2)) {
                            d_maxTheta2 += M_PI;
                            d_minTheta2 += M_PI;
                            d_minPhi2 = M_PI - d_maxPhi1;
                            d_maxPhi2 = M_PI / 2;
                            d_maxPhi1 = M_PI / 2;
                        }
                        // evaluate all four (for edge flags); combine with the reference's precedence
                        bool r1t = angle_in_range(theta, d_minTheta1, d_maxTheta1, out.edge);
                        bool r1p = angle_in_range(phi, d_minPhi1, d_maxPhi1, out.edge);
                        bool r2t = angle_in_range(theta, d_minTheta2, d_maxTheta2, out.edge);
                        bool r2p = angle_in_range(phi, d_minPhi2, d_maxPhi2, out.edge);
                        if ((r1t && r1p) || (r2t && r2p)) {
                            if (received_root == 2)
                                received_root = i;
                            else if (t[received_root] > t[i])
                                received_root = i;
                        }
                    }
                }
                if (received_root < 2) {
                    prd.end = true;
                    unsigned int i = received_root;
                    D3 endPoint;
                    endPoint.x = po.x + t[i] * dir.x;
                    endPoint.y = po.y + t[i] * dir.y;
                    endPoint.z = po.z + t[i] * dir.z;
                    D3 RxRange;
                    if ((prd.reflDepth == 0) && (prd.refrDepth == 0)) {
                        RxRange = sub(endPoint, L.origin);
                        if (length(RxRange) >= SCENE_EPS) {
                            prd.power = 1 / (4 * M_PI * 4 * M_PI * (magsq(RxRange)));
                            prd.doppler = 0;
                            prd.rayLength += t[i];
                            prd.received = Rx_i;
                            captures++;
                        }
                    } else {
                        RxRange = sub(endPoint, rd3(prd.prevHitPoint));
                        if (length(RxRange) >= SCENE_EPS_R) {
                            prd.power *= 1 / ((magsq(RxRange)) * 4 * M_PI * 4 * M_PI);
                            prd.rayLength += t[i];
                            prd.received = Rx_i;
                            captures++;
                        }
                    }
                }
            }
        }
        if (captures > 1) out.multi++;
    }

    // Earth, ray_tracer.cu:438-477
    if (prd.end == false) {
        const D3 po = rd3(prd.prevHitPoint);
        double d_earthRadius = RTS_EARTH_RADIUS;
        double A = (dir.x) * (dir.x) + (dir.y) * (dir.y) + (dir.z) * (dir.z);
        double B = 2 * (po.x * dir.x + po.y * dir.y + po.z * dir.z);
        double C = po.x * po.x + po.y * po.y + po.z * po.z - d_earthRadius * d_earthRadius;
        double discriminant = B * B - 4 * A * C;
        double t[2] = {0, 0};
        if (discriminant > 0.f) {
            discriminant = sqrt(discriminant);
            t[0] = (-B - discriminant) / (2 * A);
            t[1] = (-B + discriminant) / (2 * A);
            for (int i = 0; i < 2; i++) {
                if ((t[i] >= 0) && (prd.rayLength > 0)) {
                    prd.end = true;
                    prd.rayLength += t[i];
                }
            }
        }
    }
}

static void trace(const Launch &L, const RayF &ray, rts_ray_record &prd, uint64_t rayIndex, uint32_t col, RayOut &out,
                  bool use_bvh)
{
    out.segments++;
    const D3 o = rd3(prd.prevHitPoint), d = rd3(prd.rayDirection);
    Hit h = use_bvh ? closest_bvh(*L.scene, ray, o, d, out.edge) : closest_brute(*L.scene, ray, o, d, out.edge);
    if (h.tri >= 0) {
        out.hits++;
        if (out.tp && col < L.W) {
            uint64_t slot = prd.maxRayIndex / L.R3;
            out.tp[(slot * out.row_stride) * L.W + col] = (int32_t)h.tri;
        }
        closest_hit(L, ray, prd, h, rayIndex, col, out, use_bvh);
    } else {
        miss(L, prd, out);
    }
}

// ray_tracer.cu:144-255 for one launch index
static void ray_generation(const Launch &L, uint64_t rayIndex, RayOut &out, bool use_bvh)
{
    const uint64_t nxy = (uint64_t)L.nx * L.ny;
    const uint32_t iz = (uint32_t)(rayIndex / nxy);
    const uint32_t iy = (uint32_t)((rayIndex % nxy) / L.nx);
    const uint32_t ix = (uint32_t)(rayIndex % L.nx);
    D3 rayDir_d3 = primary_direction(L, ix, iy, iz);

    F3 rayDir_f3 = normalise_float3(rayDir_d3.x, rayDir_d3.y, rayDir_d3.z);
    RayF ray{F3{(float)L.origin.x, (float)L.origin.y, (float)L.origin.z}, rayDir_f3, SCENE_EPS,
             optix_recalled::RT_DEFAULT_MAX};

    rts_ray_record prd;
    memset(&prd, 0, sizeof(prd));
    prd.reflDepth = 0;
    prd.refrDepth = 0;
    prd.maxRayIndex = 0;
    prd.rayLength = 0;
    wd3(prd.rayDirection, rayDir_d3);
    wd3(prd.firstHitPoint, d3(0.f, 0.f, 0.f));
    wd3(prd.prevHitPoint, L.origin);
    prd.refrIndex[0] = 1; prd.refrIndex[1] = 1;
    prd.power = 0;
    prd.doppler = 0;
    prd.received = -1;
    prd.end = false;

    // :227-240 output slot init
    if (out.res) {
        for (uint32_t i = 0; i < L.M; i++) {
            rts_ray_record &r = out.res[(uint64_t)i * out.res_stride];
            memset(&r, 0, sizeof(r));
            r.refrIndex[0] = 1; r.refrIndex[1] = 1;
            r.received = -1;
            r.end = false;
        }
    }

    trace(L, ray, prd, rayIndex, 0, out, use_bvh);

    if (out.res) write_back(out.res[0], prd);
}

// ------------------------------------------------------------------------------------------
struct BinKey {
    int32_t rx;
    int32_t path[RTS_MAX_DEPTH];
    bool operator<(const BinKey &o) const
    {
        if (rx != o.rx) return rx < o.rx;
        for (uint32_t i = 0; i < RTS_MAX_DEPTH; i++)
            if (path[i] != o.path[i]) return path[i] < o.path[i];
        return false;
    }
};
struct BinAcc { double n = 0, sp = 0, sd = 0, sph = 0, sdop = 0; uint64_t min_slot = UINT64_MAX; bool direct = false; };

static void shard_bounds(const rts_pulse *p, uint64_t R3, uint64_t &b, uint64_t &e, uint64_t &stride)
{
    b = p->ray_begin;
    e = p->ray_count ? std::min(R3, p->ray_begin + p->ray_count) : R3;
    if (b > e) b = e;
    stride = p->ray_stride ? p->ray_stride : 1;
}

} // namespace orc

using namespace orc;

extern "C" int orc_sizes_for(const rts_pulse *p, orc_sizes *out)
{
    if (!p || !out) return -1;
    uint32_t rMax = p->max_refr > 0 ? 2u : 0u;
    out->rays = (uint64_t)p->nx * p->ny * p->nz;
    out->slots = rMax == 2 ? 1 + (p->max_refl + 1) + 1 : 1;
    out->ray_total = out->rays * out->slots;
    out->depth_total = p->max_refl + rMax;
    out->tri_cols = p->max_refl + 3;
    out->_pad = 0;
    return 0;
}

extern "C" void orc_rx_sphere_from_desc(const rts_rx_desc *desc, rts_rx_sphere *out)
{
    // ray_tracer.cpp:894-918 (float trig on the double angles, as written)
    double h_Rx_azimuth = desc->azimuth;
    double h_Rx_elevation = desc->elevation;
    const double r = desc->radius;
    out->centre[0] = desc->position[0] + (r * cosf(h_Rx_elevation) * cosf(h_Rx_azimuth));
    out->centre[1] = desc->position[1] + (r * cosf(h_Rx_elevation) * sinf(h_Rx_azimuth));
    out->centre[2] = desc->position[2] + (r * sinf(h_Rx_elevation));
    h_Rx_azimuth = atan2f((desc->position[1] - out->centre[1]), (desc->position[0] - out->centre[0]));
    h_Rx_elevation = atan2f((desc->position[2] - out->centre[2]),
                            sqrt((desc->position[0] - out->centre[0]) * (desc->position[0] - out->centre[0]) +
                                 (desc->position[1] - out->centre[1]) * (desc->position[1] - out->centre[1])));
    out->radius = r;
    out->min_theta = h_Rx_azimuth - desc->theta_span / 2;
    out->max_theta = h_Rx_azimuth + desc->theta_span / 2;
    out->min_phi = h_Rx_elevation - desc->phi_span / 2;
    out->max_phi = h_Rx_elevation + desc->phi_span / 2;
}

// compact: the arrays hold the shard's rays only — ray k of the shard (launch index ray_begin + k*ray_stride) has its
// slot s at index k + s*n_shard, rows likewise, edge flag at k — instead of launch-indexed arrays of the whole grid
static int trace_impl(const rts_target_mesh *targets, uint32_t n_targets, const rts_pulse *pulse, int use_bvh,
                      rts_ray_record *results, int32_t *targ_intersect, double *rcs_angle, int32_t *tri_path,
                      uint8_t *edge_flags, rts_stats *stats, bool compact)
{
    if (!pulse || (!targets && n_targets)) return -1;
    if (pulse->max_refl + (pulse->max_refr ? 2u : 0u) > RTS_MAX_DEPTH) return -2;
    Scene scene;
    build_scene(scene, targets, n_targets);
    if (use_bvh && scene.total_tris) build_bvh(scene);
    const bool bvh = use_bvh && scene.has_bvh;
    Launch L;
    setup_launch(L, scene, pulse);
    uint64_t b, e, stride;
    shard_bounds(pulse, L.R3, b, e, stride);
    const int64_t nIter = (int64_t)((e - b + stride - 1) / stride);
    const uint64_t span = compact ? (uint64_t)nIter : L.R3;     // distance between a ray's result slots
    const uint64_t ray_total = span * L.M;

    // host-side defaults (ray_tracer.cpp:854-868)
    if (targ_intersect)
        for (uint64_t i = 0; i < ray_total * L.D; i++) targ_intersect[i] = -1;
    if (rcs_angle)
        for (uint64_t i = 0; i < ray_total * L.D * 2; i++) rcs_angle[i] = -1000000;
    if (tri_path)
        for (uint64_t i = 0; i < ray_total * L.W; i++) tri_path[i] = -1;
    if (results) memset(results, 0, sizeof(rts_ray_record) * ray_total);

    uint64_t segs = 0, hits = 0, shaded = 0, refr = 0, multi = 0, edges = 0, nrays = 0, captured = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : segs, hits, shaded, refr, multi, edges, nrays, captured)
    for (int64_t it = 0; it < nIter; it++) {
        const uint64_t rayIndex = b + (uint64_t)it * stride;
        const uint64_t at = compact ? (uint64_t)it : rayIndex;
        RayOut out;
        rts_ray_record local[RTS_MAX_DEPTH + 4];
        out.res = results ? results + at : local;
        out.res_stride = results ? span : 1;
        out.ti = targ_intersect ? targ_intersect + at * L.D : nullptr;
        out.rcs = rcs_angle ? rcs_angle + at * L.D * 2 : nullptr;
        out.tp = tri_path ? tri_path + at * L.W : nullptr;
        out.row_stride = span;
        out.edge = 0;
        out.segments = out.hits = out.shaded = out.refracted = out.multi = 0;
        ray_generation(L, rayIndex, out, bvh);
        for (uint32_t k = 0; k < L.M; k++) captured += out.res[(uint64_t)k * out.res_stride].received >= 0;
        if (edge_flags) edge_flags[at] = out.edge;
        segs += out.segments; hits += out.hits; shaded += out.shaded; refr += out.refracted; multi += out.multi;
        edges += (out.edge & ORC_EDGE_TRI) ? 1 : 0;
        nrays++;
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->primary_rays = nrays; stats->segments = segs; stats->hits = hits; stats->shaded_hits = shaded;
        stats->captured = captured; stats->multi_captured = multi; stats->edge_rays = edges; stats->refracted = refr;
    }
    return 0;
}

extern "C" int orc_trace(const rts_target_mesh *targets, uint32_t n_targets, const rts_pulse *pulse, int use_bvh,
                         rts_ray_record *results, int32_t *targ_intersect, double *rcs_angle, int32_t *tri_path,
                         uint8_t *edge_flags, rts_stats *stats)
{
    return trace_impl(targets, n_targets, pulse, use_bvh, results, targ_intersect, rcs_angle, tri_path, edge_flags, stats, false);
}

extern "C" int orc_trace_shard(const rts_target_mesh *targets, uint32_t n_targets, const rts_pulse *pulse, int use_bvh,
                               rts_ray_record *results, int32_t *targ_intersect, double *rcs_angle, int32_t *tri_path,
                               uint8_t *edge_flags, rts_stats *stats)
{
    return trace_impl(targets, n_targets, pulse, use_bvh, results, targ_intersect, rcs_angle, tri_path, edge_flags, stats, true);
}

extern "C" int orc_trace_bins(const rts_target_mesh *targets, uint32_t n_targets, const rts_pulse *pulse, int use_bvh,
                              rts_bin *bins, uint32_t cap, uint32_t *n_bins, rts_stats *stats)
{
    if (!pulse || (!targets && n_targets)) return -1;
    if (pulse->max_refl + (pulse->max_refr ? 2u : 0u) > RTS_MAX_DEPTH) return -2;
    const auto t_start = std::chrono::steady_clock::now();
    Scene scene;
    build_scene(scene, targets, n_targets);
    if (use_bvh && scene.total_tris) build_bvh(scene);
    const auto t_built = std::chrono::steady_clock::now();
    const bool bvh = use_bvh && scene.has_bvh;
    Launch L;
    setup_launch(L, scene, pulse);
    uint64_t b, e, stride;
    shard_bounds(pulse, L.R3, b, e, stride);
    const int64_t nIter = (int64_t)((e - b + stride - 1) / stride);
    const double cspeed = pulse->cspeed, carrier = pulse->carrier;
    const double Wl = cspeed / carrier; // ray_tracer.cpp:815

    std::map<BinKey, BinAcc> total;
    uint64_t segs = 0, hits = 0, shaded = 0, refr = 0, multi = 0, edges = 0, nrays = 0, captured = 0;
#pragma omp parallel
    {
        std::map<BinKey, BinAcc> mine;
        std::vector<rts_ray_record> res(L.M);
        std::vector<int32_t> ti((size_t)L.M * std::max(1u, L.D));
#pragma omp for schedule(dynamic, 1024) reduction(+ : segs, hits, shaded, refr, multi, edges, nrays, captured)
        for (int64_t it = 0; it < nIter; it++) {
            const uint64_t rayIndex = b + (uint64_t)it * stride;
            std::fill(ti.begin(), ti.end(), -1);
            RayOut out;
            out.res = res.data(); out.ti = ti.data(); out.rcs = nullptr; out.tp = nullptr; out.res_stride = 1; out.row_stride = 1;
            out.edge = 0;
            out.segments = out.hits = out.shaded = out.refracted = out.multi = 0;
            ray_generation(L, rayIndex, out, bvh);
            for (uint32_t k = 0; k < L.M; k++) {
                rts_ray_record r = res[k];
                if (r.received < 0) continue;
                captured++;
                // host post-process, ray_tracer.cpp:1190-1258; Target::GetRCS -> pulse->targ_rcs[k] (NULL = 1),
                // Transmitter/Receiver::GetGain -> pulse->gain_tx / gain_rx (0 = 1)
                BinKey key;
                key.rx = r.received;
                for (uint32_t c = 0; c < RTS_MAX_DEPTH; c++) key.path[c] = -1;
                for (uint32_t c = 0; c < L.D; c++) {
                    int targ_k = ti[(size_t)k * L.D + c];
                    key.path[c] = targ_k;
                    if (targ_k >= 0) {
                        const double targRCS = pulse->targ_rcs ? pulse->targ_rcs[targ_k] : 1.0; // :1226
                        r.power *= targRCS;                                                     // :1228
                    }
                }
                const double Gt = pulse->gain_tx != 0 ? pulse->gain_tx : 1.0, Gr = pulse->gain_rx != 0 ? pulse->gain_rx : 1.0;
                r.power *= (Wl * Wl * Gt * Gr);
                double Vr = r.doppler / 2;
                r.doppler = carrier * (((1 + Vr / cspeed) / (1 - Vr / cspeed)) - 1);
                // aggregation.cu:59-65
                double delay = (r.rayLength) / cspeed;
                double phase = -fmod(delay * 2 * M_PI * carrier, 2 * M_PI);
                BinAcc &a = mine[key];
                a.n += 1;
                a.sp += sqrt(r.power);
                a.sd += delay;
                a.sph += phase;
                a.sdop += r.doppler;
                a.min_slot = std::min(a.min_slot, rayIndex + (uint64_t)k * L.R3);
                if (r.reflDepth == 0 && r.refrDepth == 0) a.direct = true;
            }
            segs += out.segments; hits += out.hits; shaded += out.shaded; refr += out.refracted; multi += out.multi;
            edges += (out.edge & ORC_EDGE_TRI) ? 1 : 0;
            nrays++;
        }
#pragma omp critical
        for (auto &kv : mine) {
            BinAcc &a = total[kv.first];
            a.n += kv.second.n; a.sp += kv.second.sp; a.sd += kv.second.sd; a.sph += kv.second.sph; a.sdop += kv.second.sdop;
            a.min_slot = std::min(a.min_slot, kv.second.min_slot);
            a.direct = a.direct || kv.second.direct;
        }
    }
    // per-receiver totals for the direct-ray rule (aggregation.cu:56)
    std::map<int32_t, BinAcc> rxTot;
    for (auto &kv : total) {
        BinAcc &t = rxTot[kv.first.rx];
        t.n += kv.second.n; t.sp += kv.second.sp; t.sd += kv.second.sd; t.sph += kv.second.sph; t.sdop += kv.second.sdop;
        t.min_slot = std::min(t.min_slot, kv.second.min_slot);
    }
    uint32_t n = 0;
    for (auto &kv : total) {
        if (n < cap && bins) {
            rts_bin &o = bins[n];
            memset(&o, 0, sizeof(o));
            o.rx = kv.first.rx;
            memcpy(o.path, kv.first.path, sizeof(o.path));
            const BinAcc &a = kv.second.direct ? rxTot[kv.first.rx] : kv.second;
            o.direct = kv.second.direct ? 1 : 0;
            o.npath = a.n; o.sum_sqrt_power = a.sp; o.sum_delay = a.sd; o.sum_phase = a.sph; o.sum_doppler = a.sdop;
            o.min_slot = a.min_slot;
            o.own_min_slot = kv.second.min_slot;
            // aggregation.cu:86-92
            if (a.n > 0) {
                o.power = pow(a.sp / a.n, 2);
                o.delay = a.sd / a.n;
                o.phase = a.sph / a.n;
                o.doppler = a.sdop / a.n;
            }
        }
        n++;
    }
    if (n_bins) *n_bins = n;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->primary_rays = nrays; stats->segments = segs; stats->hits = hits; stats->shaded_hits = shaded;
        stats->captured = captured; stats->multi_captured = multi; stats->edge_rays = edges; stats->refracted = refr;
        stats->n_bins = n;
        const auto t_end = std::chrono::steady_clock::now();
        stats->ms_update = std::chrono::duration<float, std::milli>(t_built - t_start).count();   // oracle BVH build
        stats->ms_trace = std::chrono::duration<float, std::milli>(t_end - t_built).count();      // trace + aggregate
        stats->ms_total = stats->ms_update + stats->ms_trace;
    }
    return 0;
}

extern "C" int orc_num_threads(void) { return omp_get_max_threads(); }
// torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; a timed oracle run states its thread count itself
extern "C" void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
extern "C" const char *orc_version(void) { return "rts-oracle 1 (scalar C++/OpenMP restatement; test infrastructure)"; }
