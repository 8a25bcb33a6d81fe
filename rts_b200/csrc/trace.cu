// trace.cu — the wavefront bounce kernel (sm_100a): ray generation, BVH traversal with the fp64
// ray/triangle test, closest-hit shading (reflect / refract / Doppler / path rows), miss
// (receiver capture) and fused receiver-bin accumulation.
//
// One launch = one bounce wave over a queue of ray states (struct of arrays, 8-byte coalesced
// accesses).  Persistent CTAs; each warp claims 32 queue entries at a time from a global work
// counter.  Surviving rays and refracted children are appended to the next wave's queue with
// warp-aggregated atomics (cooperative_groups::coalesced_threads), which is the compaction step.
//
// Reference semantics reproduced here (file:line in /root/reference):
//   ray_generation    ray_tracer.cu:144-255       intersect  triangle_mesh.cu:121-200
//   closest_hit       normal_shader.cu:128-340    miss       ray_tracer.cu:260-478
//   host post-process ray_tracer.cpp:1190-1258    myKernel1  aggregation.cu:56-69 (per-ray terms)
// The recursion of the reference (rtTrace inside closest_hit) becomes: the reflected ray continues
// in place as the same chain; a refracted ray is a new chain (result slot +1) pushed to the queue.
#include "engine.h"
#include "dev_math.cuh"
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cstring>
#include <algorithm>

namespace cg = cooperative_groups;

namespace {

constexpr float RT_DEFAULT_MAX_F = 1.e27f;
constexpr double EDGE_EPS = 1e-9;

struct Ray {
    double ox, oy, oz, dx, dy, dz, len, pw, dop, fx, fy, fz, n0, n1;
    unsigned long long key;
    uint32_t ray, meta;
};

__device__ __forceinline__ uint32_t m_refl(uint32_t m) { return m & 0xffu; }
__device__ __forceinline__ uint32_t m_refr(uint32_t m) { return (m >> 8) & 3u; }
__device__ __forceinline__ uint32_t m_slot(uint32_t m) { return (m >> 10) & 15u; }
__device__ __forceinline__ bool m_end(uint32_t m) { return (m >> 14) & 1u; }
__device__ __forceinline__ bool m_primary(uint32_t m) { return (m >> 15) & 1u; }
__device__ __forceinline__ uint32_t m_col(uint32_t m) { return (m >> 16) & 15u; }
// bit 20: first reflection of a primary ray whose primary hit is the kept static one (coherent.cuh)
constexpr uint32_t M_COH = 1u << 20;
__device__ __forceinline__ uint32_t m_make(uint32_t refl, uint32_t refr, uint32_t slot, bool end, bool prim, uint32_t col)
{
    return (refl & 0xffu) | ((refr & 3u) << 8) | ((slot & 15u) << 10) | ((end ? 1u : 0u) << 14) | ((prim ? 1u : 0u) << 15) |
           ((col & 15u) << 16);
}

// The state is loaded in two parts: what traversal needs (origin, direction, flags) before it, the
// rest (length, power, Doppler, first hit, indices, path key) only after it — that keeps ~20 registers
// free during the traversal loop (occupancy is the lever: see profiles/).
// Queue columns are streamed once per wave: evict-first loads and stores (ld.global.cs / st.global.cs) keep them from
// pushing the BVH nodes and triangle records out of L2.
__device__ __forceinline__ void load_ray_geom(const RayQueue &q, unsigned long long i, Ray &r)
{
    r.ox = __ldcs(q.f[F_OX] + i); r.oy = __ldcs(q.f[F_OY] + i); r.oz = __ldcs(q.f[F_OZ] + i);
    r.dx = __ldcs(q.f[F_DX] + i); r.dy = __ldcs(q.f[F_DY] + i); r.dz = __ldcs(q.f[F_DZ] + i);
    r.meta = __ldcs(q.meta + i);
}
// The first hit point is only ever reported through the records, the refractive indices only matter with a
// refraction budget: those five columns are skipped otherwise (88 instead of 128 bytes per queued ray).
__device__ __forceinline__ void load_ray_rest(const RayQueue &q, unsigned long long i, Ray &r, bool records, bool refr)
{
    r.len = __ldcs(q.f[F_LEN] + i); r.pw = __ldcs(q.f[F_PW] + i); r.dop = __ldcs(q.f[F_DOP] + i);
    if (records) { r.fx = __ldcs(q.f[F_FX] + i); r.fy = __ldcs(q.f[F_FY] + i); r.fz = __ldcs(q.f[F_FZ] + i); }
    else { r.fx = 0; r.fy = 0; r.fz = 0; }
    if (refr) { r.n0 = __ldcs(q.f[F_N0] + i); r.n1 = __ldcs(q.f[F_N1] + i); }
    else { r.n0 = 1; r.n1 = 1; }
    r.key = __ldcs(q.key + i); r.ray = __ldcs(q.ray + i);
}
__device__ __forceinline__ void store_ray(const RayQueue &q, unsigned long long i, const Ray &r, bool records, bool refr)
{
    __stcs(q.f[F_OX] + i, r.ox); __stcs(q.f[F_OY] + i, r.oy); __stcs(q.f[F_OZ] + i, r.oz);
    __stcs(q.f[F_DX] + i, r.dx); __stcs(q.f[F_DY] + i, r.dy); __stcs(q.f[F_DZ] + i, r.dz);
    __stcs(q.f[F_LEN] + i, r.len); __stcs(q.f[F_PW] + i, r.pw); __stcs(q.f[F_DOP] + i, r.dop);
    if (records) { __stcs(q.f[F_FX] + i, r.fx); __stcs(q.f[F_FY] + i, r.fy); __stcs(q.f[F_FZ] + i, r.fz); }
    if (refr) { __stcs(q.f[F_N0] + i, r.n0); __stcs(q.f[F_N1] + i, r.n1); }
    __stcs(q.key + i, r.key); __stcs(q.ray + i, r.ray); __stcs(q.meta + i, r.meta);
}

// Append one ray per calling thread to the next wave's queue: one atomic per converged group.
__device__ __forceinline__ void push_ray(const WaveParams &P, const Ray &r, unsigned &overflow)
{
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(P.out_count, (unsigned long long)g.size());
    base = g.shfl(base, 0);
    const unsigned long long i = base + g.thread_rank();
    if (i < P.out_capacity) store_ray(P.out, i, r, (P.flags & RTS_OUT_RECORDS) != 0 || P.keep_first != 0, P.rMax != 0);
    else overflow++;
}

// The same from the far end of the queue (refracted children when the queue is filled from both ends).
__device__ __forceinline__ void push_ray_back(const WaveParams &P, const Ray &r, unsigned &overflow)
{
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(P.out_back, (unsigned long long)g.size());
    base = g.shfl(base, 0);
    const unsigned long long i = base + g.thread_rank();
    if (i < P.out_capacity) store_ray(P.out, P.out_capacity - 1ull - i, r, (P.flags & RTS_OUT_RECORDS) != 0 || P.keep_first != 0, P.rMax != 0);
    else overflow++;
}
// Slot of entry i of the incoming wave (see WaveParams::in_back).
__device__ __forceinline__ unsigned queue_slot(const WaveParams &P, unsigned i, unsigned n_front)
{
    return i < n_front ? i : (unsigned)P.out_capacity - 1u - (i - n_front);
}

// ray_tracer.cu:158-204 with the launch-invariant trigonometry hoisted to the host (WaveParams).
__device__ __forceinline__ d3 primary_direction(const WaveParams &P, uint32_t ix, uint32_t iy, uint32_t iz)
{
    if (P.single_ray) return mk3(P.boresight[0], P.boresight[1], P.boresight[2]);
    d3 d;
    d.x = P.nx > 1 ? P.beamStart[0] + P.slope[0] * (ix) : P.beamStart[0];
    d.y = P.ny > 1 ? P.beamStart[1] + P.slope[1] * (iy) : P.beamStart[1];
    d.z = P.nz > 1 ? P.beamStart[2] + P.slope[2] * (iz) : P.beamStart[2];
    d = normalised3_shared_rcp(d);
    d3 r = mk3(0, 0, 0);
    r.x += P.Rot[0] * d.x + P.Rot[1] * d.y + P.Rot[2] * d.z;
    r.y += P.Rot[3] * d.x + P.Rot[4] * d.y + P.Rot[5] * d.z;
    r.z += P.Rot[6] * d.x + P.Rot[7] * d.y + P.Rot[8] * d.z;
    d = normalised3_shared_rcp(r);
    r = mk3(0, 0, 0);
    r.x += P.Rot1[0] * d.x + P.Rot1[1] * d.y + P.Rot1[2] * d.z;
    r.y += P.Rot1[3] * d.x + P.Rot1[4] * d.y + P.Rot1[5] * d.z;
    r.z += P.Rot1[6] * d.x + P.Rot1[7] * d.y + P.Rot1[8] * d.z;
    return r;
}

struct Tri { d3 p0, p1, p2; uint32_t id, target; };

__device__ __forceinline__ Tri load_tri(const TriRec *rec, int pos)
{
    const double2 *s = reinterpret_cast<const double2 *>(rec + pos);
    const double2 a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2), d = __ldg(s + 3), e = __ldg(s + 4);
    Tri t;
    t.p0 = mk3(a.x, a.y, b.x); t.p1 = mk3(b.y, c.x, c.y); t.p2 = mk3(d.x, d.y, e.x);
    const unsigned long long ids = (unsigned long long)__double_as_longlong(e.y);
    t.id = (uint32_t)ids; t.target = (uint32_t)(ids >> 32);
    return t;
}

// triangle_mesh.cu:121-137
__device__ __forceinline__ bool tri_test(const Tri &T, const d3 &o, const d3 &dir, double tmin_d, double tmax_d, d3 &n,
                                         double &t, double &beta, double &gamma)
{
    const d3 e0 = T.p1 - T.p0;
    const d3 e1 = T.p0 - T.p2;
    n = cross3(e1, e0);
    const d3 e2 = (1 / dot3(n, dir)) * (T.p0 - o);
    const d3 i = cross3(dir, e2);
    beta = dot3(i, e1);
    gamma = dot3(i, e0);
    t = dot3(n, e2);
    return ((t < tmax_d) & (t > tmin_d) & (beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1));
}

// The same test, same operations and rounding, evaluated t-first so that candidates outside
// (tmin, tmax) skip the barycentric part; accept/reject and t are bit-identical to tri_test.
__device__ __forceinline__ bool tri_accept(const Tri &T, const d3 &o, const d3 &dir, double tmin_d, double tmax_d, double &t)
{
    const d3 e0 = T.p1 - T.p0;
    const d3 e1 = T.p0 - T.p2;
    const d3 n = cross3(e1, e0);
    const d3 e2 = (1 / dot3(n, dir)) * (T.p0 - o);
    t = dot3(n, e2);
    if (!((t < tmax_d) & (t > tmin_d))) return false;
    const d3 i = cross3(dir, e2);
    const double beta = dot3(i, e1);
    const double gamma = dot3(i, e0);
    return ((beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1));
}

struct HitRec { int pos; float t; uint32_t id; };
#ifndef RTS_THIN_CLAIM
#define RTS_THIN_CLAIM 4u      // rays a warp of a thin (chained) wave claims at a time
#endif

// Packed fp32 pairs (sm_100a FFMA2 / FADD2): one instruction, two lanes, each lane rounded like the scalar op.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 r, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// BVH traversal.  Boxes are tested in fp32, centre/half-extent form, with a rigorous error bound so
// that the test is conservative with respect to the exact slab interval of the true fp64 ray (o, d).
// Per axis a, for a box [c-h, c+h] (c, h exact fp32 values; the box contains the reference's leaf box):
//   exact      near* = (c - o)/d - h/|d|          far* = (c - o)/d + h/|d|
//   computed   T = fma(c, inv, -oi)               inv = rcp.approx(fl32(d)), oi = fl32(fl32(o) * inv)   (all fp32)
//              H = fma(h, |inv|, E)               E   = 2^-20 * (S + |o|) * |inv|
//              near = T - H                       far = T + H
// With S = max |coordinate| of the scene box on the axis (|c| <= S, h <= S), u = 1/|d|, and the relative errors
// |inv d - 1| <= 2^-24 + 2^-23 (narrowing + MUFU.RCP, PTX: max relative error 2^-23), |oi d/o - 1| <= 2.6 * 2^-23:
//   |T - (c-o)/d| <= 2^-23 (2.1 |c| + 3.1 |o|) u,  H >= (h u (1 - 1.6 * 2^-23) + E)(1 - 2^-24),
//   |fl(T -+ H) - (T -+ H)| <= 2^-24 (|T| + H)
// so near <= near* and far >= far* whenever E >= 2^-23 (5.2 S + 3.6 |o|) u, i.e. 0.65 * 2^-20 (S+|o|) u: a box is never
// pruned when the fp64 triangle test (triangle_mesh.cu:121-137) could accept a hit inside it.  Only the bound
// matters here, not bit-reproducibility, so fused multiply-adds are fine.  Axes with |fl32(d)| < 1e-20 are ignored
// (T = 0, H = inf: always overlapping).  The x,y lanes of one child and the z lanes of both children are packed
// fp32 pairs, matching the BvhNode word order: 12 packed instructions test both child boxes.
// Closest hit = smallest fp32 t in (tmin, inf), ties to the lowest global triangle id
// (rtPotentialIntersection semantics with a defined tie rule).
// BUDGET (k_primary_follow): a ray whose walk needs more than RTS_FOLLOW_BUDGET leaf visits / stack pops is given up
// (the return value says so; best is then meaningless) so that the caller can queue it for the next, thin launch instead
// of holding 31 finished lanes and, at the end of the launch, the whole GPU (see follow.cuh).
#ifndef RTS_FOLLOW_BUDGET
#define RTS_FOLLOW_BUDGET 48u
#endif
template <bool COUNT, bool STATIC_ONLY = false, bool BUDGET = false>
__device__ __forceinline__ bool traverse(const WaveParams &P, const d3 &o, const d3 &dir, float tmin_f, HitRec &best,
                                         unsigned &n_nodes, unsigned &n_tris, unsigned &stack_ovf)
{
    best.pos = -1; best.t = RT_DEFAULT_MAX_F; best.id = 0xffffffffu;
    if (P.n_tris == 0) return true;
    unsigned rounds = 0;
    u64 inv_xy, noi_xy, ainv_xy, e_xy, inv_zz, noi_zz, ainv_zz, e_zz;
    {
        const float oo[3] = {(float)o.x, (float)o.y, (float)o.z}, dd[3] = {(float)dir.x, (float)dir.y, (float)dir.z};
        float inv[3], noi[3], ainv[3], E[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (!(fabsf(dd[a]) >= 1e-20f)) {
                inv[a] = 0.f; noi[a] = 0.f; ainv[a] = 0.f; E[a] = CUDART_INF_F;
            } else {
                float r;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dd[a]));
                inv[a] = r;
                noi[a] = -(oo[a] * r);
                ainv[a] = fabsf(r);
                E[a] = 9.5367431640625e-07f * ((P.scene_abs[a] + fabsf(oo[a])) * ainv[a]) + 1e-30f;
            }
        }
        inv_xy = pk2(inv[0], inv[1]); noi_xy = pk2(noi[0], noi[1]); ainv_xy = pk2(ainv[0], ainv[1]); e_xy = pk2(E[0], E[1]);
        inv_zz = pk2(inv[2], inv[2]); noi_zz = pk2(noi[2], noi[2]); ainv_zz = pk2(ainv[2], ainv[2]); e_zz = pk2(E[2], E[2]);
    }
    const double tmin_d = (double)tmin_f, tmax_d = (double)RT_DEFAULT_MAX_F;
    float best_pad = CUDART_INF_F;
    // "while-while" (Aila & Laine 2009): each lane descends internal nodes until it stands on a leaf (or has
    // nothing left); the warp reconverges at the end of the inner loop, so the fp64 triangle code below runs on
    // as full a warp as the rays allow instead of being replayed for a few lanes per iteration.
    constexpr int SENT = 0x7fffffff;       // empty stack; node indices are < 2^28
    int stack[RTS_STACK_DEPTH];
    int sp = 0;
    int cur = P.root_ref;
    while (cur != SENT) {
        if (BUDGET && ++rounds > RTS_FOLLOW_BUDGET) return false;
        while ((unsigned)cur < (unsigned)SENT) {
            if (COUNT) n_nodes++;
            const ulonglong2 *np = reinterpret_cast<const ulonglong2 *>(P.nodes + cur);
            const ulonglong2 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);   // (c0xy,h0xy) (c1xy,h1xy) (czz,hzz)
            const int2 refs = __ldg(reinterpret_cast<const int2 *>(np + 3));
            const u64 T0 = fma2(q0.x, inv_xy, noi_xy), H0 = fma2(q0.y, ainv_xy, e_xy);
            const u64 T1 = fma2(q1.x, inv_xy, noi_xy), H1 = fma2(q1.y, ainv_xy, e_xy);
            const u64 Tz = fma2(q2.x, inv_zz, noi_zz), Hz = fma2(q2.y, ainv_zz, e_zz);
            float n0x, n0y, f0x, f0y, n1x, n1y, f1x, f1y, nz0, nz1, fz0, fz1;
            upk2(sub2(T0, H0), n0x, n0y); upk2(add2(T0, H0), f0x, f0y);
            upk2(sub2(T1, H1), n1x, n1y); upk2(add2(T1, H1), f1x, f1y);
            upk2(sub2(Tz, Hz), nz0, nz1); upk2(add2(Tz, Hz), fz0, fz1);
            const float tn0 = fmaxf(fmaxf(n0x, n0y), nz0), tf0 = fminf(fminf(f0x, f0y), fz0);
            const float tn1 = fmaxf(fmaxf(n1x, n1y), nz1), tf1 = fminf(fminf(f1x, f1y), fz1);
            // overlap of [tn, tf] with [0, best_pad]
            const bool h0 = fmaxf(tn0, 0.f) <= fminf(tf0, best_pad);
            const bool h1 = fmaxf(tn1, 0.f) <= fminf(tf1, best_pad);
            if (h0 & h1) {
                const bool swap = tn1 < tn0;
                const int far_ref = swap ? refs.x : refs.y;
                if (sp < RTS_STACK_DEPTH) stack[sp++] = far_ref;
                else stack_ovf++;
#ifdef RTS_PREFETCH_FAR
                if (far_ref >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.nodes + far_ref));
#endif
                cur = swap ? refs.y : refs.x;
            } else if (h0) cur = refs.x;
            else if (h1) cur = refs.y;
            else cur = sp ? stack[--sp] : SENT;
        }
        if (cur < 0) {
            const int code = ~cur;
            const int first = code >> 3, cnt = (code & 7) + 1;
            for (int k = 0; k < cnt; k++) {
                if (COUNT) n_tris++;
                const Tri T = load_tri(P.trirec, first + k);
                if (STATIC_ONLY && P.moving_flags[T.target]) continue;   // closest hit among the triangles that never move
                double t;
                if (tri_accept(T, o, dir, tmin_d, tmax_d, t)) {
                    const float tf = (float)t;
                    if (tf > tmin_f && (tf < best.t || (tf == best.t && T.id < best.id))) {
                        best.t = tf; best.pos = first + k; best.id = T.id;
                        best_pad = tf * 1.000001f;
                    }
                }
            }
            cur = sp ? stack[--sp] : SENT;
        }
    }
    return true;
}

// ---- the same traversal over the quantised 32-byte nodes (engine.h: QNode), for rays that start inside the scene ----
// A plane of a child box is x(q) = lo + q * ext / 32768 on the frame (lo, ext) of the scene (fp32 values, exact reals
// here); k_pack rounds every box outward to such planes and widens it by one more cell.  Per axis, per ray:
//     inv = rcp.approx(fl32(d))                               relative error <= 1.5 * 2^-23
//     A   = fl32(ext * inv)                                   B = fl32((lo - o) * inv - A)       ((lo - o) * inv - A in fp64)
//     t   = fma(val, A, B),  val = 1 + q / 32768              (exact: one PRMT of the node's half word with 0x3F800000)
// B is formed from the rounded A, so val * A + B = (q / 32768) * A + (lo - o) * inv: A's rounding only touches the
// first term.  With r = |lo - o| / ext and cell = ext / 32768, in units of cell * |inv|:
//     inv:  1.5 * 2^-23 * |t| <= 1.5 * 2^-8 (r + 1)     A: 2^-9     B: 2^-9 (r + 1)     fma: 2^-9 (r + 1)
// together < 0.0098 (r + 1) + 0.002 cells — 0.02 cells for an origin inside the frame (every ray of a later wave: it starts
// on a triangle), under one cell for r <= 64; beyond that (and for |d| < 1e-20) the axis is ignored: near = -inf, far =
// 3e38.  The one-cell widening of the boxes therefore makes the test conservative with respect to the exact slab
// interval of the fp64 ray, like traverse()'s.  The sign of inv picks which half word is the near plane (PRMT selector).
struct QRay {
    u64 a[3], b[3];            // (A, A), (B, B): lanes = child 0, child 1
    uint32_t sn[3], sf[3];     // PRMT selectors of the near / far plane
};
constexpr uint32_t Q_MAGIC = 0x3F800000u;

__device__ __forceinline__ void q_setup(const WaveParams &P, const d3 &o, const d3 &dir, QRay &Q)
{
    const double oo[3] = {o.x, o.y, o.z};
    const float dd[3] = {(float)dir.x, (float)dir.y, (float)dir.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float flo = __ldg(&P.qframe->lo[k]), ext = __ldg(&P.qframe->ext[k]);
        const double rel = (double)flo - oo[k];
        float A = 3.0e38f, B = 0.f;
        uint32_t sn = 0xE654u, sf = 0x7654u;          // -inf and 1.0 from the magic word alone
        if ((fabsf(dd[k]) >= 1e-20f) && (fabs(rel) <= 64.0 * (double)ext)) {
            float inv;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(dd[k]));
            A = ext * inv;
            B = (float)(rel * (double)inv - (double)A);
            sn = inv >= 0.f ? 0x7104u : 0x7324u;
            sf = inv >= 0.f ? 0x7324u : 0x7104u;
        }
        Q.a[k] = pk2(A, A); Q.b[k] = pk2(B, B);
        Q.sn[k] = sn; Q.sf[k] = sf;
    }
}

// both child boxes of node `cur`: overlap of [tn, tf] with [0, best_pad] per child, entry distances, child references
__device__ __forceinline__ void q_visit(const QNode *__restrict__ qnodes, int cur, const QRay &Q, float best_pad, bool &h0, bool &h1,
                                        float &tn0, float &tn1, int2 &refs)
{
    const uint4 *np = reinterpret_cast<const uint4 *>(qnodes + cur);
    const uint4 q0 = __ldg(np), q1 = __ldg(np + 1);      // (c0x, c0y, c0z, c1x) (c1y, c1z, ref0, ref1)
    refs = make_int2((int)q1.z, (int)q1.w);
#define QV(w, sel) __uint_as_float(__byte_perm((w), Q_MAGIC, (sel)))
    const u64 Nx = fma2(pk2(QV(q0.x, Q.sn[0]), QV(q0.w, Q.sn[0])), Q.a[0], Q.b[0]);
    const u64 Fx = fma2(pk2(QV(q0.x, Q.sf[0]), QV(q0.w, Q.sf[0])), Q.a[0], Q.b[0]);
    const u64 Ny = fma2(pk2(QV(q0.y, Q.sn[1]), QV(q1.x, Q.sn[1])), Q.a[1], Q.b[1]);
    const u64 Fy = fma2(pk2(QV(q0.y, Q.sf[1]), QV(q1.x, Q.sf[1])), Q.a[1], Q.b[1]);
    const u64 Nz = fma2(pk2(QV(q0.z, Q.sn[2]), QV(q1.y, Q.sn[2])), Q.a[2], Q.b[2]);
    const u64 Fz = fma2(pk2(QV(q0.z, Q.sf[2]), QV(q1.y, Q.sf[2])), Q.a[2], Q.b[2]);
#undef QV
    float n0x, n1x, n0y, n1y, n0z, n1z, f0x, f1x, f0y, f1y, f0z, f1z;
    upk2(Nx, n0x, n1x); upk2(Ny, n0y, n1y); upk2(Nz, n0z, n1z);
    upk2(Fx, f0x, f1x); upk2(Fy, f0y, f1y); upk2(Fz, f0z, f1z);
    tn0 = fmaxf(fmaxf(n0x, n0y), n0z); tn1 = fmaxf(fmaxf(n1x, n1y), n1z);
    const float tf0 = fminf(fminf(f0x, f0y), f0z), tf1 = fminf(fminf(f1x, f1y), f1z);
    h0 = fmaxf(tn0, 0.f) <= fminf(tf0, best_pad);
    h1 = fmaxf(tn1, 0.f) <= fminf(tf1, best_pad);
}

template <bool COUNT, bool STATIC_ONLY = false>
__device__ __forceinline__ void traverse_q(const WaveParams &P, const d3 &o, const d3 &dir, float tmin_f, HitRec &best,
                                           unsigned &n_nodes, unsigned &n_tris, unsigned &stack_ovf)
{
    best.pos = -1; best.t = RT_DEFAULT_MAX_F; best.id = 0xffffffffu;
    if (P.n_tris == 0) return;
    QRay Q;
    q_setup(P, o, dir, Q);
    const double tmin_d = (double)tmin_f, tmax_d = (double)RT_DEFAULT_MAX_F;
    float best_pad = CUDART_INF_F;
    constexpr int SENT = 0x7fffffff;
    int stack[RTS_STACK_DEPTH];
    int sp = 0;
    int cur = P.root_ref;
    while (cur != SENT) {
        while ((unsigned)cur < (unsigned)SENT) {
            if (COUNT) n_nodes++;
            bool h0, h1;
            float tn0, tn1;
            int2 refs;
            q_visit(P.qnodes, cur, Q, best_pad, h0, h1, tn0, tn1, refs);
            if (h0 & h1) {
                const bool swap = tn1 < tn0;
                if (sp < RTS_STACK_DEPTH) stack[sp++] = swap ? refs.x : refs.y;
                else stack_ovf++;
                cur = swap ? refs.y : refs.x;
            } else if (h0) cur = refs.x;
            else if (h1) cur = refs.y;
            else cur = sp ? stack[--sp] : SENT;
        }
        if (cur < 0) {
            const int code = ~cur;
            const int first = code >> 3, cnt = (code & 7) + 1;
            for (int k = 0; k < cnt; k++) {
                if (COUNT) n_tris++;
                const Tri T = load_tri(P.trirec, first + k);
                if (STATIC_ONLY && P.moving_flags[T.target]) continue;
                double t;
                if (tri_accept(T, o, dir, tmin_d, tmax_d, t)) {
                    const float tf = (float)t;
                    if (tf > tmin_f && (tf < best.t || (tf == best.t && T.id < best.id))) {
                        best.t = tf; best.pos = first + k; best.id = T.id;
                        best_pad = tf * 1.000001f;
                    }
                }
            }
            cur = sp ? stack[--sp] : SENT;
        }
    }
}

// Chain end: the record the reference writes back for this result slot
// (ray_tracer.cu:246-253 for slot 0, normal_shader.cu:272-279 for refracted slots).
template <bool RECORDS>
__device__ __forceinline__ void finish_chain(const WaveParams &P, const Ray &r, int received)
{
    if (RECORDS) {
        rts_ray_record *o = P.results + ((unsigned long long)r.ray + (unsigned long long)m_slot(r.meta) * P.R3);
        o->reflDepth = m_refl(r.meta);
        o->refrDepth = m_refr(r.meta);
        o->rayLength = r.len;
        o->firstHitPoint[0] = r.fx; o->firstHitPoint[1] = r.fy; o->firstHitPoint[2] = r.fz;
        o->prevHitPoint[0] = r.ox; o->prevHitPoint[1] = r.oy; o->prevHitPoint[2] = r.oz;
        o->power = r.pw;
        o->doppler = r.dop;
        o->received = received;
    }
}

__device__ __forceinline__ void cart_to_sph(d3 in, double &azi, double &ele) // normal_shader.cu:118-124
{
    azi = atan2(in.y, in.x);
    ele = atan2(in.z, sqrt(in.x * in.x + in.y * in.y));
}

// Per-thread event counters, packed three 21-bit fields to a register pair so they do not crowd the
// traversal loop: a thread can see at most 2^26 / 32 rays per launch (queue capacity 3 * 2^24 < 2^26 rays,
// claimed 32 at a time by its warp), so no field can overflow.
struct Local {
    unsigned long long a;        // hits | shaded << 21 | captured << 42
    unsigned long long b;        // multi | edge << 21 | refracted << 42
    unsigned long long nodes, tris;   // only touched by the COUNT instantiation
    unsigned overflow;
};
constexpr unsigned long long C_HIT = 1ull, C_SHADED = 1ull << 21, C_CAPTURED = 1ull << 42;
constexpr unsigned long long C_MULTI = 1ull, C_EDGE = 1ull << 21, C_REFRACTED = 1ull << 42;

// Bilinear sample of a tabulated callback (rts_table2d), arguments clamped to the grid.
__device__ __forceinline__ double table_lookup(const DevTable &t, double az, double el)
{
    double u = (az - t.az0) / t.az_step, v = (el - t.el0) / t.el_step;
    u = fmin(fmax(u, 0.0), (double)(t.n_az - 1u));
    v = fmin(fmax(v, 0.0), (double)(t.n_el - 1u));
    const uint32_t i0 = min((uint32_t)u, t.n_az - 1u), j0 = min((uint32_t)v, t.n_el - 1u);
    const uint32_t i1 = min(i0 + 1u, t.n_az - 1u), j1 = min(j0 + 1u, t.n_el - 1u);
    const double f = u - (double)i0, g = v - (double)j0;
    const double v00 = t.values[(size_t)i0 * t.n_el + j0], v01 = t.values[(size_t)i0 * t.n_el + j1];
    const double v10 = t.values[(size_t)i1 * t.n_el + j0], v11 = t.values[(size_t)i1 * t.n_el + j1];
    return (v00 * (1 - f) + v10 * f) * (1 - g) + (v01 * (1 - f) + v11 * f) * g;
}
// SVec3(Vec3) of the host simulator: azimuth = atan2(y, x), elevation = asin(z / length); a zero vector keeps (0, 0)
__device__ __forceinline__ void svec3_angles(double x, double y, double z, double &az, double &el)
{
    const double len = sqrt(x * x + y * y + z * z);
    az = 0; el = 0;
    if (len != 0) { el = asin(z / len); az = atan2(y, x); }
}

// Tabulated Target::GetRCS for one hop (RTS_TABLES): the summed in/out angles of normal_shader.cu:320-326, then the
// target's table (or its scalar).  Out of line: pulses without tables must not pay registers for it.
__device__ __noinline__ double rcs_hop_factor(const WaveParams &P, uint32_t targ, const d3 k0, const d3 k1)
{
    const DevTable tab = P.rcs_tab[targ];
    if (!tab.n_az) return P.t_rcs ? P.t_rcs[targ] : 1.0;
    double a0, e0, a1, e1;
    cart_to_sph(k0, a0, e0);
    cart_to_sph(mk3(-k1.x, -k1.y, -k1.z), a1, e1);
    return table_lookup(tab, a0 + a1, e0 + e1);
}

// normal_shader.cu:128-340
// Returns true when `chain` is set and the reflected ray is to be followed at once by the caller (r holds it)
// instead of being queued for the next wave.
template <bool RECORDS, bool TABLES = false>
__device__ __forceinline__ bool shade(const WaveParams &P, Ray &r, const HitRec &h, Local &L, bool chain, uint32_t coh = 0)
{
    const uint32_t dMax = P.dMax, rMax = P.rMax;
    uint32_t reflDepth = m_refl(r.meta), refrDepth = m_refr(r.meta);
    const uint32_t slot = m_slot(r.meta), col = m_col(r.meta);
    bool end = m_end(r.meta);
    const bool primary = m_primary(r.meta);
    const Tri T = load_tri(P.trirec, h.pos);
    const unsigned long long row = (unsigned long long)r.ray + (unsigned long long)slot * P.R3;

    if (RECORDS && col < P.W) P.tri_path[row * P.W + col] = (int32_t)T.id;

    // guard, :134 — an absorbed hit changes nothing
    if (!((end == false) && ((refrDepth < rMax) || (reflDepth < (dMax - 1))))) {
        finish_chain<RECORDS>(P, r, -1);
        return false;
    }
    L.a += C_SHADED;

    // the winner's attributes, recomputed with the same arithmetic as during traversal
    d3 o = mk3(r.ox, r.oy, r.oz), dir = mk3(r.dx, r.dy, r.dz);
    d3 n;
    double tt, beta, gamma;
    tri_test(T, o, dir, 0.0, 0.0, n, tt, beta, gamma);
    if (fmin(fmin(beta, gamma), 1 - beta - gamma) < EDGE_EPS) L.b += C_EDGE;
    const uint32_t targ = T.target;
    d3 normal;
    if (P.interpolate) { // triangle_mesh.cu:177-190
        const uint32_t noff = P.t_norm_off[targ];
        if (P.t_per_face[targ]) {
            const double *nn = P.world_normals + 3 * (size_t)(noff + (T.id - P.t_tri_off[targ]));
            normal = mk3(nn[0], nn[1], nn[2]);
        } else {
            const uint32_t *vi = P.tris + 3 * (size_t)T.id;
            const double *a0 = P.world_normals + 3 * (size_t)(noff + vi[0]);
            const double *a1 = P.world_normals + 3 * (size_t)(noff + vi[1]);
            const double *a2 = P.world_normals + 3 * (size_t)(noff + vi[2]);
            normal = mk3(a1[0] * beta + a2[0] * gamma + a0[0] * (1.0f - beta - gamma),
                         a1[1] * beta + a2[1] * gamma + a0[1] * (1.0f - beta - gamma),
                         a1[2] * beta + a2[2] * gamma + a0[2] * (1.0f - beta - gamma));
        }
        normal = normalised3(normal);
    } else {
        normal = normalised3(n);
    }
    const double reflCoeff = P.t_refl[targ], refrIndex = P.t_refr[targ];
    const float hit_t = h.t;

    // :140-146 path row
    if (refrDepth != 1) {
        const uint32_t x = reflDepth + refrDepth;
        if (x < (rMax + dMax - 1)) {
            if (RECORDS) P.targ_intersect[row * P.D + x] = (int)targ;
            r.key += (unsigned long long)(targ + 1) * P.powB[x];
        }
    }

    // :148-152
    d3 hitPoint;
    hitPoint.x = o.x + (double)hit_t * dir.x;
    hitPoint.y = o.y + (double)hit_t * dir.y;
    hitPoint.z = o.z + (double)hit_t * dir.z;
    r.len += hit_t;

    // :158-173 spreading loss
    if ((reflDepth == 0) && (refrDepth == 0)) {
        r.fx = hitPoint.x; r.fy = hitPoint.y; r.fz = hitPoint.z;
        const d3 TxRange = hitPoint - mk3(P.origin[0], P.origin[1], P.origin[2]);
        if (len3(TxRange) >= SCENE_EPS) r.pw = 1 / ((magsq3(TxRange)) * 4 * M_PI);
        else end = true;
    } else {
        const d3 TargRange = hitPoint - o;
        if (len3(TargRange) >= SCENE_EPS_R) r.pw *= 1 / ((magsq3(TargRange)) * 4 * M_PI);
        else end = true;
    }

    // :176
    r.ox = hitPoint.x; r.oy = hitPoint.y; r.oz = hitPoint.z;

    // the fp32 direction of the ray being shaded (ray.direction): primary rays carry
    // normalise_float3(fp64 dir) (ray_tracer.cu:208), later rays the fp32 direction itself
    f3 rdir;
    if (primary) rdir = normalise_float3(dir);
    else { rdir.x = (float)dir.x; rdir.y = (float)dir.y; rdir.z = (float)dir.z; }
    const f3 nf = normalise_float3(normal);
    const d3 V = mk3(P.t_vel[3 * targ], P.t_vel[3 * targ + 1], P.t_vel[3 * targ + 2]);
    // The unit vectors k0, k1 (normal_shader.cu:251-253, 302-304) feed only the Doppler term V.(k1-k0)
    // and the RCS angles.  For a target at rest V.(k1-k0) is exactly +0 and doppler += 0 leaves the value
    // unchanged, so the two fp64 normalisations are skipped unless the angles are wanted.
    const bool want_rcs = RECORDS && !(P.flags & RTS_NO_RCS_ANGLES);
    // tabulated Target::GetRCS (RTS_TABLES, fused bins): the factor of this hop is folded into the power where the reference
    // writes the hop's angles (:320-326) — the host loop of ray_tracer.cpp:1221-1231 multiplies the very same factors later
    const bool fold_rcs = TABLES && !RECORDS && P.rcs_tab != nullptr;   // TABLES: separate instantiations, so that pulses without tables pay nothing
    const bool need_k = want_rcs || fold_rcs || (V.x != 0.0) || (V.y != 0.0) || (V.z != 0.0);

    // :191-194
    const double pr_n0 = r.n1; // prd_refr.refrIndex.x = prd_refr.refrIndex.y

    // :198-282 refraction: a new chain in the next result slot
    if ((fabs(reflCoeff) != 1.00000f) && (refrDepth < rMax) && (reflDepth == 0)) {
        const double pr_n1 = (pr_n0 == 1) ? refrIndex : 1.0;
        const float ratio = (float)(pr_n1 / pr_n0);
        f3 nd;
        if (optix_refract(nd, rdir, nf, ratio)) {
            L.b += C_REFRACTED;
            const uint32_t cslot = slot + 1;
            const unsigned long long crow = (unsigned long long)r.ray + (unsigned long long)cslot * P.R3;
            Ray c = r;
            c.n0 = pr_n0; c.n1 = pr_n1;
            if ((refrDepth == 0) && (cslot == 1)) { // :221-239 pre-fill
                if (RECORDS) {
                    for (uint32_t i = 0; i < P.D; i++) P.targ_intersect[crow * P.D + i] = (int)targ;
                    for (uint32_t j = 0; j < dMax; j++)
                        for (uint32_t i = 0; i < (j + 2) && i < P.D; i++)
                            P.targ_intersect[((unsigned long long)r.ray + (unsigned long long)(j + 2) * P.R3) * P.D + i] = (int)targ;
                }
                c.key = (unsigned long long)(targ + 1) * P.key_all;
            } else {
                // exit chain: its row was pre-filled with the first-hit target in columns 0 and 1
                const unsigned long long first_digit = r.key % P.powB[1];
                c.key = first_digit * (1ull + P.powB[1]);
            }
            if ((reflDepth + 1) < dMax) c.pw *= (1 - fabs(reflCoeff)); // :245-246
            const uint32_t c_refr = refrDepth + 1;                       // :247
            c.dx = (double)nd.x; c.dy = (double)nd.y; c.dz = (double)nd.z;
            if (need_k) {
                const d3 k0 = normalised3(dir);                          // :251-256
                const d3 k1 = normalised3(mk3(c.dx, c.dy, c.dz));
                c.dop += dot3(V, k1 - k0);
            if (want_rcs) { // :259-265
                const uint32_t x = reflDepth + (c_refr - 1);
                if (x < P.D) {
                    double a0, e0, a1, e1;
                    cart_to_sph(k0, a0, e0);
                    cart_to_sph(mk3(-k1.x, -k1.y, -k1.z), a1, e1);
                    P.rcs_angle[(crow * P.D + x) * 2] = a0 + a1;
                    P.rcs_angle[(crow * P.D + x) * 2 + 1] = e0 + e1;
                }
            }
            }
            c.meta = m_make(reflDepth, c_refr, cslot, end, false, col + 1);
            if (P.out_back) push_ray_back(P, c, L.overflow); // :268
            else push_ray(P, c, L.overflow);
        }
    }

    // :286-290
    reflDepth++;
    r.n1 = pr_n0;
    r.n0 = pr_n0;

    // :293-333 reflection: the same chain continues
    if (reflDepth < dMax) {
        const f3 nd = optix_reflect(rdir, nf);
        r.pw *= reflCoeff;
        r.dx = (double)nd.x; r.dy = (double)nd.y; r.dz = (double)nd.z;
        if (need_k) {
        const d3 k0 = normalised3(dir);
        const d3 k1 = normalised3(mk3(r.dx, r.dy, r.dz));
        r.dop += dot3(V, k1 - k0);
        if (want_rcs) { // :320-326
            const uint32_t x = (reflDepth - 1) + refrDepth;
            if (x < P.D) {
                double a0, e0, a1, e1;
                cart_to_sph(k0, a0, e0);
                cart_to_sph(mk3(-k1.x, -k1.y, -k1.z), a1, e1);
                P.rcs_angle[(row * P.D + x) * 2] = a0 + a1;
                P.rcs_angle[(row * P.D + x) * 2 + 1] = e0 + e1;
            }
        }
        if (fold_rcs && ((reflDepth - 1) + refrDepth) < P.D) r.pw *= rcs_hop_factor(P, targ, k0, k1);
        }
        r.meta = m_make(reflDepth, refrDepth, slot, end, false, col + 1) | coh;
        if (chain) return true;
        push_ray(P, r, L.overflow); // :332
    } else {
        r.meta = m_make(reflDepth, refrDepth, slot, end, false, col + 1);
        finish_chain<RECORDS>(P, r, -1);
    }
    return false;
}

// ray_tracer.cu:260-478.  Returns the receiver index or -1.
template <bool RECORDS>
__device__ __forceinline__ int miss(const WaveParams &P, Ray &r, Local &L)
{
    const uint32_t reflDepth = m_refl(r.meta), refrDepth = m_refr(r.meta);
    bool end = m_end(r.meta);
    int received = -1;
    const d3 dir = mk3(r.dx, r.dy, r.dz);
    const d3 po = mk3(r.ox, r.oy, r.oz);
    if (end == false) {
        double A, B, C, discriminant;
        double t[2] = {0, 0};
        unsigned captures = 0;
        for (unsigned int Rx_i = 0; Rx_i < P.n_rx; Rx_i++) {
            const RxDev rx = P.rx[Rx_i];
            A = (dir.x) * (dir.x) + (dir.y) * (dir.y) + (dir.z) * (dir.z);
            B = 2 * (((po.x - rx.cx) * dir.x) + ((po.y - rx.cy) * dir.y) + ((po.z - rx.cz) * dir.z));
            C = po.x * po.x + po.y * po.y + po.z * po.z + (rx.cx * rx.cx) + (rx.cy * rx.cy) + (rx.cz * rx.cz) -
                2 * ((rx.cx * po.x) + (rx.cy * po.y) + (rx.cz * po.z)) - rx.radius * rx.radius;
            discriminant = B * B - 4 * A * C;
            if (discriminant > 0.f) {
                discriminant = sqrt(discriminant);
                t[0] = (-B - discriminant) / (2 * A);
                t[1] = (-B + discriminant) / (2 * A);
                unsigned int received_root = 2;
                for (int i = 0; i < 2; i++) {
                    if ((t[i] >= 0) && ((r.len + t[i]) > SCENE_EPS) && ((r.len + t[i]) > SCENE_EPS_R)) {
                        d3 endPoint;
                        endPoint.x = po.x + t[i] * dir.x;
                        endPoint.y = po.y + t[i] * dir.y;
                        endPoint.z = po.z + t[i] * dir.z;
                        double theta = atan2f((endPoint.y - rx.cy), (endPoint.x - rx.cx));
                        double phi = atan2f(endPoint.z - rx.cz, sqrt(((endPoint.y - rx.cy) * (endPoint.y - rx.cy)) +
                                                                     ((endPoint.x - rx.cx) * (endPoint.x - rx.cx))));
                        if ((phi < -M_PI / 2)) { theta += M_PI; phi = -M_PI - phi; }
                        if ((phi > M_PI / 2)) { theta += M_PI; phi = M_PI - phi; }
                        double maxTheta1 = rx.max_theta, minTheta1 = rx.min_theta;
                        double maxTheta2 = maxTheta1, minTheta2 = minTheta1;
                        double maxPhi1 = rx.max_phi, minPhi1 = rx.min_phi;
                        double maxPhi2 = maxPhi1, minPhi2 = minPhi1;
                        if ((minPhi1 < -M_PI / 2)) {
                            maxTheta2 += M_PI; minTheta2 += M_PI;
                            maxPhi2 = -M_PI - minPhi1; minPhi2 = -M_PI / 2; minPhi1 = -M_PI / 2;
                        }
                        if ((maxPhi1 > M_PI / 2)) {
                            maxTheta2 += M_PI; minTheta2 += M_PI;
                            minPhi2 = M_PI - maxPhi1; maxPhi2 = M_PI / 2; maxPhi1 = M_PI / 2;
                        }
                        if (((angle_in_range(theta, minTheta1, maxTheta1)) && (angle_in_range(phi, minPhi1, maxPhi1))) ||
                            ((angle_in_range(theta, minTheta2, maxTheta2)) && (angle_in_range(phi, minPhi2, maxPhi2)))) {
                            if (received_root == 2) received_root = i;
                            else if (t[received_root] > t[i]) received_root = i;
                        }
                    }
                }
                if (received_root < 2) {
                    end = true;
                    const unsigned int i = received_root;
                    d3 endPoint;
                    endPoint.x = po.x + t[i] * dir.x;
                    endPoint.y = po.y + t[i] * dir.y;
                    endPoint.z = po.z + t[i] * dir.z;
                    if ((reflDepth == 0) && (refrDepth == 0)) {
                        const d3 RxRange = endPoint - mk3(P.origin[0], P.origin[1], P.origin[2]);
                        if (len3(RxRange) >= SCENE_EPS) {
                            r.pw = 1 / (4 * M_PI * 4 * M_PI * (magsq3(RxRange)));
                            r.dop = 0;
                            r.len += t[i];
                            received = (int)Rx_i;
                            captures++;
                        }
                    } else {
                        const d3 RxRange = endPoint - po;
                        if (len3(RxRange) >= SCENE_EPS_R) {
                            r.pw *= 1 / ((magsq3(RxRange)) * 4 * M_PI * 4 * M_PI);
                            r.len += t[i];
                            received = (int)Rx_i;
                            captures++;
                        }
                    }
                }
            }
        }
        if (captures > 1) L.b += C_MULTI;
    }
    // Earth (ray_tracer.cu:438-477): only observable through rayLength of unreceived records
    if (RECORDS && end == false) {
        const double R = RTS_EARTH_RADIUS;
        const double A = (dir.x) * (dir.x) + (dir.y) * (dir.y) + (dir.z) * (dir.z);
        const double B = 2 * (po.x * dir.x + po.y * dir.y + po.z * dir.z);
        const double C = po.x * po.x + po.y * po.y + po.z * po.z - R * R;
        double discriminant = B * B - 4 * A * C;
        if (discriminant > 0.f) {
            discriminant = sqrt(discriminant);
            const double t0 = (-B - discriminant) / (2 * A), t1 = (-B + discriminant) / (2 * A);
            if ((t0 >= 0) && (r.len > 0)) { end = true; r.len += t0; }
            if ((t1 >= 0) && (r.len > 0)) { end = true; r.len += t1; }
        }
    }
    finish_chain<RECORDS>(P, r, received);
    return received;
}

// Wl^2 * Gt * Gr of a captured ray with tabulated antenna patterns (RTS_TABLES):
// Gt = trans->GetGain(transvec, GetRotation(time_t)), Gr = recv->GetGain(recvvec, GetRotation(delay + time_t))
// (ray_tracer.cpp:1201-1235, 1247); an antenna without a table keeps its scalar gain.  Out of line like rcs_hop_factor.
__device__ __noinline__ double antenna_factor(const WaveParams &P, const Ray &r, int received)
{
    double Gt = P.gain_tx_scalar, Gr = P.gain_rx_scalar;
    const bool direct = (m_refl(r.meta) == 0) && (m_refr(r.meta) == 0);
    const DevAntenna rxa = P.ant_rx[received];
    double az, el;
    if (P.ant_tx && P.ant_tx->gain.n_az) {
        if (direct) svec3_angles(P.origin[0] - rxa.pos[0], P.origin[1] - rxa.pos[1], P.origin[2] - rxa.pos[2], az, el);   // :1206
        else svec3_angles(r.fx - P.origin[0], r.fy - P.origin[1], r.fz - P.origin[2], az, el);                              // :1210
        Gt = table_lookup(P.ant_tx->gain, az - P.ant_tx->bore_az, el - P.ant_tx->bore_el);
    }
    if (rxa.gain.n_az) {
        if (direct) svec3_angles(rxa.pos[0] - P.origin[0], rxa.pos[1] - P.origin[1], rxa.pos[2] - P.origin[2], az, el);   // :1207
        else svec3_angles(r.ox - rxa.pos[0], r.oy - rxa.pos[1], r.oz - rxa.pos[2], az, el);                                 // :1211
        const double d = r.len / P.cspeed;                                                                                  // :1218
        Gr = table_lookup(rxa.gain, az - (rxa.bore_az + rxa.rate_az * d), el - (rxa.bore_el + rxa.rate_el * d));
    }
    return P.wl2 * Gt * Gr;
}

// Fused host post-process (per-target scalar RCS, constant Gt/Gr; ray_tracer.cpp:1219-1253) and the per-ray terms of
// myKernel1 (aggregation.cu:59-69), summed into the (receiver, path) bin.  Lanes of a converged
// group that hit the same bin are reduced with shuffles first, so one group issues one set of
// fp64 atomics per distinct bin.
// Shared-memory pre-reduction of the bins (the north star's "shared-memory pre-reduction ... before any global atomics"):
// when the dense table is small (P.smem_bins = n_bins <= RTS_SMEM_BINS) every CTA accumulates into its own copy — the
// lanes of a warp are combined with shuffles first (below), the warps of the CTA with shared-memory atomics — and adds
// the copy to the global table once, when its warps have run out of rays: one set of global atomics per CTA and touched
// bin instead of one per warp and capture.  Large tables (the benchmark's 5,832 bins) stay with global atomics: a copy
// per CTA would take L1 away from the BVH nodes.
extern __shared__ unsigned long long s_bins_raw[];
__device__ __forceinline__ void bins_smem_init(const WaveParams &P)
{
    if (!P.smem_bins) return;
    const unsigned n = P.smem_bins;
    for (unsigned i = threadIdx.x; i < n * 6u; i += blockDim.x) s_bins_raw[i] = i < n * 5u ? 0ull : 0x7f7f7f7f7f7f7f7full;
    __syncthreads();
}
__device__ __forceinline__ void bins_smem_flush(const WaveParams &P)
{
    if (!P.smem_bins) return;
    __syncthreads();
    const unsigned n = P.smem_bins;
    const double *s_sum = reinterpret_cast<const double *>(s_bins_raw);
    for (unsigned b = threadIdx.x; b < n; b += blockDim.x) {
        if (s_sum[b * 5u] != 0.0) {                      // npath > 0: the bin was touched by this CTA
#pragma unroll
            for (int k = 0; k < 5; k++) atomicAdd(P.bin_sums + (size_t)b * 5 + k, s_sum[b * 5u + k]);
            atomicMin(P.bin_mins + b, s_bins_raw[n * 5u + b]);
        }
    }
}

template <bool TABLES = false>
__device__ __forceinline__ void accumulate_bin(const WaveParams &P, const Ray &r, int received)
{
    double pw = r.pw;
    if (P.t_rcs && !(TABLES && P.rcs_tab)) { // ray_tracer.cpp:1219-1230: one factor per path-row entry >= 0, in column order
        unsigned long long k = r.key;
        for (uint32_t c = 0; c < P.D; c++) {
            const uint32_t digit = (uint32_t)(k % P.B);
            k /= P.B;
            if (digit) pw *= P.t_rcs[digit - 1];
        }
    }
    pw *= (TABLES && P.ant_rx) ? antenna_factor(P, r, received) : P.wl2gain;
    const double Vr = r.dop / 2;
    const double dopHz = P.carrier * (((1 + Vr / P.cspeed) / (1 - Vr / P.cspeed)) - 1);
    const double delay = (r.len) / P.cspeed;
    const double phase = -fmod(delay * 2 * M_PI * P.carrier, 2 * M_PI);
    const double amp = sqrt(pw);
    const unsigned long long bin = (unsigned long long)received * P.powB[P.D] + r.key;
    const unsigned long long slot = (unsigned long long)r.ray + (unsigned long long)m_slot(r.meta) * P.R3;
    if (!P.hash_keys && bin >= P.n_bins) return;
    cg::coalesced_group g = cg::coalesced_threads();
    auto part = cg::labeled_partition(g, bin);
    const double s_n = cg::reduce(part, 1.0, cg::plus<double>());
    const double s_a = cg::reduce(part, amp, cg::plus<double>());
    const double s_d = cg::reduce(part, delay, cg::plus<double>());
    const double s_p = cg::reduce(part, phase, cg::plus<double>());
    const double s_f = cg::reduce(part, dopHz, cg::plus<double>());
    const unsigned long long s_m = cg::reduce(part, slot, cg::less<unsigned long long>());
    if (part.thread_rank() == 0) {
        unsigned long long at = bin;
        if (P.hash_keys) {   // sparse table: find or claim the bin's slot (linear probing; a claimed slot is listed once)
            const unsigned long long mask = P.n_bins - 1ull;
            unsigned long long h = bin * 0x9E3779B97F4A7C15ull;
            h ^= h >> 29;
            at = h & mask;
            unsigned long long probes = 0;
            for (;; at = (at + 1ull) & mask) {
                unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(P.hash_keys + at);
                if (cur == ~0ull) {
                    cur = atomicCAS(P.hash_keys + at, ~0ull, bin);
                    if (cur == ~0ull) { P.hash_used[atomicAdd(P.hash_count, 1u)] = (uint32_t)at; break; }
                }
                if (cur == bin) break;
                if (++probes > mask) { atomicAdd(&P.counters->overflow, 1ull); return; }   // table full: reported as RTS_ERR_CAPACITY
            }
        }
        if (P.smem_bins) {          // this CTA's copy (dense table: at == bin < n_bins)
            double *b = reinterpret_cast<double *>(s_bins_raw) + at * 5;
            atomicAdd(b + 0, s_n);
            atomicAdd(b + 1, s_a);
            atomicAdd(b + 2, s_d);
            atomicAdd(b + 3, s_p);
            atomicAdd(b + 4, s_f);
            atomicMin(s_bins_raw + (size_t)P.smem_bins * 5 + at, s_m);
            return;
        }
        double *b = P.bin_sums + at * 5;
        atomicAdd(b + 0, s_n);
        atomicAdd(b + 1, s_a);
        atomicAdd(b + 2, s_d);
        atomicAdd(b + 3, s_p);
        atomicAdd(b + 4, s_f);
        atomicMin(P.bin_mins + at, s_m);
    }
}

#include "raster.cuh"
#include "coherent.cuh"
#include "split.cuh"
#include "follow.cuh"

template <bool PRIMARY, bool RECORDS, bool COUNT, bool CHAIN, bool TABLES = false>
__global__ void __launch_bounds__(RTS_WAVE_BLOCK, PRIMARY ? RTS_WAVE_MIN_BLOCKS_PRIMARY : RTS_WAVE_MIN_BLOCKS) k_wave(const __grid_constant__ WaveParams P)
{
    if (PRIMARY && P.raster_ctl && raster_on(P)) return;   // the projected primary wave (raster.cuh) did this batch
    const unsigned lane = threadIdx.x & 31u;
    // queue capacity and batch size are < 2^26, so 32-bit indices and a 32-bit work counter suffice
    // a queue that overflowed in the previous wave holds out_capacity entries, however far its counters ran on
    const unsigned n_front = PRIMARY ? (unsigned)P.n_primary : (unsigned)min(*P.in_count, P.out_capacity);
    const unsigned n_all = n_front + ((!PRIMARY && P.in_back) ? (unsigned)min(*P.in_back, P.out_capacity - n_front) : 0u);
    const unsigned n_in = (!PRIMARY && P.todo_list) ? (unsigned)*P.todo_count : n_all;
    if (!PRIMARY && P.split_on && n_in >= P.split_below && P.n_tris != 0) return;   // k_traverse + k_shade_wave did this wave (split.cuh)
    unsigned *work = reinterpret_cast<unsigned *>(P.work_counter);
    Local L = {0, 0, 0, 0, 0};
    bins_smem_init(P);
    // Thin late waves (a few thousand rays whose latency, not throughput, sets the launch time) follow their
    // reflections in place instead of paying one more launch per bounce; refracted children are still queued.
    // A separate instantiation (CHAIN, used from the third wave on) so the bulk waves keep their register budget.
    const bool chain = CHAIN && !PRIMARY && n_in < P.chain_below;
    unsigned chained = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(P.wave_segs + P.wave_index, (unsigned long long)n_all);
        atomicAdd(&P.counters->segments, (unsigned long long)n_all);   // one closest-hit query per queue entry
    }
    // A thin wave is latency-bound with most issue slots idle: its warps claim RTS_THIN_CLAIM rays at a time instead of
    // 32, so eight times as many warps are in flight and a warp waits for the longest of 4 chains, not of 32 (0.145 -> 0.107 ms for the 25 k rays of the benchmark's third wave; 16 / 8 rays per claim: 0.132 / 0.115 ms).
    const unsigned claim = chain ? RTS_THIN_CLAIM : 32u;
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(work, claim);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_in) break;
        unsigned idx = base + lane;
        if (lane >= claim || idx >= n_in) continue;
        Ray r;
        unsigned long long rayIndex = 0;
        if (PRIMARY) {
            // shard-local index -> launch index; whole 8x4 tiles of the (y,z) plane per warp where the shape allows
            unsigned long long g = P.batch_base + idx;
            if (g < P.swz_limit) {
                const unsigned tile = (unsigned)(g >> 5), w = (unsigned)g & 31u, tpr = P.swz_w >> 3;
                const unsigned ty = tile / tpr, tx = tile - ty * tpr;
                g = (unsigned long long)(ty * 4u + (w >> 3)) * P.swz_w + tx * 8u + (w & 7u);
            }
            rayIndex = P.ray_begin + g * P.ray_stride;
            const unsigned long long nxy = (unsigned long long)P.nx * P.ny;
            const uint32_t iz = (uint32_t)(rayIndex / nxy);
            const uint32_t iy = (uint32_t)((rayIndex % nxy) / P.nx);
            const uint32_t ix = (uint32_t)(rayIndex % P.nx);
            const d3 d = primary_direction(P, ix, iy, iz);
            r.ox = P.origin[0]; r.oy = P.origin[1]; r.oz = P.origin[2];
            r.dx = d.x; r.dy = d.y; r.dz = d.z;
            r.meta = m_make(0, 0, 0, false, true, 0);
        } else {
            if (P.todo_list) idx = P.todo_list[idx];         // the rays k_wave1_kept left for a full trace (coherent.cuh): slots
            else idx = queue_slot(P, idx, n_front);
            load_ray_geom(P.in, idx, r);
            r.meta &= ~M_COH;
        }
        for (bool first = true;; first = false) {
            HitRec h;
            unsigned nn = 0, nt = 0;
            // SCENE_EPS == SCENE_EPS_R (ray_tracer.h:9-10).  Later waves start on a triangle, i.e. inside the frame of the
            // quantised nodes; a primary ray starts at the transmitter, wherever that is: fp32 nodes
            // (the cooperative finish of straggler rays, follow.cuh, was tried here too: on the ship scene a large share of
            // the rays inside the closed dielectric hull exceeds any small budget, and walking them one per warp cost 23 %)
            if (PRIMARY || !RTS_QNODES) traverse<COUNT>(P, mk3(r.ox, r.oy, r.oz), mk3(r.dx, r.dy, r.dz), SCENE_EPS, h, nn, nt, L.overflow);
            else traverse_q<COUNT>(P, mk3(r.ox, r.oy, r.oz), mk3(r.dx, r.dy, r.dz), SCENE_EPS, h, nn, nt, L.overflow);
            if (COUNT) { L.nodes += nn; L.tris += nt; }
            if (PRIMARY) {
                r.len = 0; r.pw = 0; r.dop = 0; r.fx = 0; r.fy = 0; r.fz = 0; r.n0 = 1; r.n1 = 1;
                r.key = 0; r.ray = (uint32_t)rayIndex;
            } else if (first) {
                load_ray_rest(P.in, idx, r, RECORDS || P.keep_first != 0, P.rMax != 0);
            }
            bool follow = false;
            if (h.pos >= 0) {
                L.a += C_HIT;
                follow = shade<RECORDS, TABLES>(P, r, h, L, chain);
            } else {
                const int received = miss<RECORDS>(P, r, L);
                if (received >= 0) {
                    L.a += C_CAPTURED;
                    if (P.flags & RTS_OUT_BINS) accumulate_bin<TABLES>(P, r, received);
                }
            }
            if (PRIMARY || !CHAIN || !follow) break;
            chained++;
        }
    }
    bins_smem_flush(P);
    if (!PRIMARY && CHAIN) {   // segments traced in place belong to this launch
        const unsigned x = __reduce_add_sync(0xffffffffu, chained);
        if (lane == 0 && x) {
            atomicAdd(P.wave_segs + P.wave_index, (unsigned long long)x);
            atomicAdd(&P.counters->segments, (unsigned long long)x);
        }
    }
    // counters: unpack, warp reduce (redux.sync), one atomic per warp per non-zero counter
    unsigned long long *c = reinterpret_cast<unsigned long long *>(P.counters);
    const unsigned f[7] = {(unsigned)(L.a & 0x1fffff), (unsigned)((L.a >> 21) & 0x1fffff), (unsigned)(L.a >> 42),
                           (unsigned)(L.b & 0x1fffff), (unsigned)((L.b >> 21) & 0x1fffff), (unsigned)(L.b >> 42), L.overflow};
    // Counters layout: segments, hits, shaded, captured, multi, edge, refracted, nodes, tris, overflow
    const int slot[7] = {1, 2, 3, 4, 5, 6, 9};
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const unsigned x = __reduce_add_sync(0xffffffffu, f[k]);
        if (lane == 0 && x) atomicAdd(c + slot[k], (unsigned long long)x);
    }
    if (COUNT) {
        unsigned long long v[2] = {L.nodes, L.tris};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            unsigned long long x = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0 && x) atomicAdd(c + 7 + k, x);
        }
    }
}

} // namespace

int trace_wave_grid(rts_engine *e)
{
    if (e->wave_grid) return e->wave_grid;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wave<true, false, false, false>, RTS_WAVE_BLOCK, 0);
    e->wave_grid_primary = e->num_sms * (occ > 0 ? occ : 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wave<false, false, false, false>, RTS_WAVE_BLOCK, 0);
    e->wave_grid = e->num_sms * (occ > 0 ? occ : 1);
    return e->wave_grid;
}

static size_t bins_smem_bytes(const WaveParams &p) { return (size_t)p.smem_bins * 48u; }

template <bool PRIMARY, bool CHAIN>
static void launch_variant(int grid, cudaStream_t st, const WaveParams &p, bool records, bool count)
{
    if (p.rcs_tab || p.ant_rx) {   // RTS_TABLES pulses (fused bins, no records, no node counting): their own instantiation
        k_wave<PRIMARY, false, false, CHAIN, true><<<grid, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
        return;
    }
    if (records) {
        if (count) k_wave<PRIMARY, true, true, CHAIN><<<grid, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
        else k_wave<PRIMARY, true, false, CHAIN><<<grid, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
    } else {
        if (count) k_wave<PRIMARY, false, true, CHAIN><<<grid, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
        else k_wave<PRIMARY, false, false, CHAIN><<<grid, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
    }
}

// per-warp scratch of the cooperative straggler traversal, for kernels of up to `ctas` CTAs
static int ensure_coop(rts_engine *e, int ctas)
{
    if (e->coop_ctas >= ctas) return RTS_OK;
    if (e->d_coop_stacks) { cudaFree(e->d_coop_stacks); e->d_coop_stacks = nullptr; e->coop_ctas = 0; }
    RTS_CUDA(cudaMalloc(&e->d_coop_stacks, COOP_WARP_BYTES * (size_t)ctas * (RTS_WAVE_BLOCK / 32)));
    e->coop_ctas = ctas;
    return RTS_OK;
}

int trace_launch_wave(rts_engine *e, const WaveParams &p_in, bool primary, bool records)
{
    trace_wave_grid(e);
    bvh_join(e);
    const int grid = primary ? e->wave_grid_primary : e->wave_grid;   // persistent: resident CTAs per SM x SMs
    const WaveParams &p = p_in;
    const bool count = (p.flags & RTS_COUNT_NODES) != 0;
    if (primary) launch_variant<true, false>(grid, e->stream, p, records, count);
    else if (p.wave_index >= (e->followed ? 1u : 2u) && p.chain_below) launch_variant<false, true>(grid, e->stream, p, records, count);
    else launch_variant<false, false>(grid, e->stream, p, records, count);
    RTS_CUDA(cudaGetLastError());
    e->launches++;
    return RTS_OK;
}

// The two-kernel form of a later wave (split.cuh), enqueued ahead of the fused kernel; each of the three decides from
// the wave's size on the device whether it is the one to run.
int trace_launch_split(rts_engine *e, WaveParams &p, bool records)
{
    if (e->trav_alloc < e->q_capacity) {
        if (e->d_trav_hits) cudaFree(e->d_trav_hits);
        e->d_trav_hits = nullptr; e->trav_alloc = 0;
        RTS_CUDA(cudaMalloc(&e->d_trav_hits, sizeof(unsigned long long) * e->q_capacity));
        e->trav_alloc = e->q_capacity;
    }
    if (!e->trav_grid) {
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_traverse<false>, RTS_WAVE_BLOCK, 0);
        e->trav_grid = e->num_sms * (occ > 0 ? occ : 1);
    }
    bvh_join(e);
    p.trav_hits = e->d_trav_hits;
    p.split_on = 1;
    // p.split_below: set by the caller (knob; default 2^18 rays)
    p.split_keep_all = records ? 1u : 0u;
    const bool timed = p.wave_index == 1 && e->split_ev[0];   // the second wave's two kernels, timed apart (rts_get_split_profile)
    if (timed) cudaEventRecord(e->split_ev[0], e->stream);
    if (p.flags & RTS_COUNT_NODES) k_traverse<true><<<e->trav_grid, RTS_WAVE_BLOCK, 0, e->stream>>>(p);
    else k_traverse<false><<<e->trav_grid, RTS_WAVE_BLOCK, 0, e->stream>>>(p);
    if (timed) cudaEventRecord(e->split_ev[1], e->stream);
    if (records) k_shade_wave<true><<<e->num_sms * RTS_SHADE_MIN_BLOCKS, RTS_WAVE_BLOCK, bins_smem_bytes(p), e->stream>>>(p);
    else k_shade_wave<false><<<e->num_sms * RTS_SHADE_MIN_BLOCKS, RTS_WAVE_BLOCK, bins_smem_bytes(p), e->stream>>>(p);
    if (timed) { cudaEventRecord(e->split_ev[2], e->stream); e->split_timed = true; }
    RTS_CUDA(cudaGetLastError());
    e->launches += 2;
    return RTS_OK;
}

#define RTS_RASTER_ITEM_CAP (1u << 20)

int trace_raster_alloc(rts_engine *e, uint64_t batch)
{
    if (e->raster_alloc >= batch) return RTS_OK;
    void **ptrs[] = {(void **)&e->d_dirs, (void **)&e->d_hits, (void **)&e->d_hits_static, &e->d_raster_ctl, &e->d_raster_ctl_static,
                     &e->d_raster_items};
    for (void **p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }   // (cudaFree waits for the device: nothing of the side streams is in flight)
    e->raster_alloc = 0;
    RTS_CUDA(cudaMalloc(&e->d_dirs, sizeof(double) * 3 * batch));
    RTS_CUDA(cudaMalloc(&e->d_hits, sizeof(unsigned long long) * batch));
    e->dirs_free_valid = false;
    RTS_CUDA(cudaMalloc(&e->d_hits_static, sizeof(unsigned long long) * batch));
    if (e->d_w1_static) { cudaFree(e->d_w1_static); e->d_w1_static = nullptr; }
    RTS_CUDA(cudaMalloc(&e->d_w1_static, sizeof(unsigned long long) * batch));
    e->w1_valid = false;
    RTS_CUDA(cudaMalloc(&e->d_raster_ctl, sizeof(RasterCtl)));
    RTS_CUDA(cudaMalloc(&e->d_raster_ctl_static, sizeof(RasterCtl)));
    RTS_CUDA(cudaMalloc(&e->d_raster_items, sizeof(RasterItem) * (size_t)RTS_RASTER_ITEM_CAP));
    e->raster_alloc = batch;
    e->dirs_valid = false;
    e->static_valid = false;
    return RTS_OK;
}

// One footprint pass (setup + small + large footprints) over the triangles p selects.
static void launch_footprints(rts_engine *e, const WaveParams &p, unsigned count, cudaStream_t stream = nullptr)
{
    const cudaStream_t fs = stream ? stream : e->stream;
    const unsigned bs = 128;
    const unsigned blocks = (count + bs - 1) / bs;
    if (!blocks) return;
    k_raster_small<<<blocks, bs, 0, fs>>>(p);       // footprints + guard + the small footprints' candidates
    k_raster_big<<<e->num_sms * 8, bs, 0, fs>>>(p);
    e->launches += 2;
}

int trace_launch_raster(rts_engine *e, WaveParams &p, bool records, bool single_batch)
{
    cudaStream_t st = e->stream;
    p.raster_ctl = (RasterCtl *)e->d_raster_ctl;
    p.raster_items = (RasterItem *)e->d_raster_items;
    p.raster_item_cap = RTS_RASTER_ITEM_CAP;
    // a warp walks a chunk 32 candidates at a time: 2048 per chunk in large launches, but a small launch whose few
    // triangles fill the image (flat plate, 65 k rays: 64 chunks = 64 warps x 64 rounds, 0.06 ms) needs more, shorter ones
    p.raster_chunk = 128u;
    while (p.raster_chunk < RTS_RASTER_CHUNK && (uint64_t)p.raster_chunk * 2048u < p.n_primary) p.raster_chunk *= 2u;
    p.leaf_of_tri = e->d_leaf_of_tri;
    p.raster_list = nullptr; p.raster_list_count = 0; p.raster_skip = nullptr; p.raster_static = nullptr;
    RTS_CUDA(cudaMemsetAsync(e->d_raster_ctl, 0, sizeof(RasterCtl), st));
    // The directions depend on the launch geometry only (Tx boresight and span, grid, shard, batch) — not on the scene:
    // when a pulse repeats the previous pulse's launch (a staring transmitter), the buffer is still valid.
    DirsKey key;
    memset(&key, 0, sizeof(key));
    key.nx = p.nx; key.ny = p.ny; key.nz = p.nz; key.single = p.single_ray;
    key.begin = p.ray_begin; key.stride = p.ray_stride; key.base = p.batch_base; key.n = p.n_primary;
    memcpy(key.c, p.origin, sizeof(double) * 3); memcpy(key.c + 3, p.beamStart, sizeof(double) * 3);
    memcpy(key.c + 6, p.slope, sizeof(double) * 3); memcpy(key.c + 9, p.Rot, sizeof(double) * 9);
    memcpy(key.c + 18, p.Rot1, sizeof(double) * 9); memcpy(key.c + 27, p.boresight, sizeof(double) * 3);
    cudaEvent_t *tl = nullptr;       // knob debug_timeline
    if (e->knobs.debug_timeline && e->tl_n < 16 && (p.flags & RTS_NO_REUSE)) {
        tl = e->tl_ev[e->tl_n++];
        for (int k = 0; k < 7; k++) if (!tl[k]) cudaEventCreate(&tl[k]);
    }
    const bool reuse = !(p.flags & RTS_NO_REUSE);
    const bool same_launch = reuse && e->dirs_valid && memcmp(&key, &e->dirs_key, sizeof(key)) == 0;
    for (int a = 0; a < 3; a++) p.dirs[a] = e->d_dirs + (size_t)a * e->raster_alloc;
    p.hits = e->d_hits;
    // head_open: work of this batch is in flight on side_dirs that the engine's stream has not been told to wait for yet
    bool head_open = false;
    auto join_head = [&]() {
        if (!head_open) return;
        cudaEventRecord(e->ev_dirs_done, e->side_dirs);
        cudaStreamWaitEvent(st, e->ev_dirs_done, 0);
        head_open = false;
    };
    if (!same_launch) {
        // on its own stream (engine.h: side_dirs): the two buffers are free once the previous shading pass and the BVH
        // primary wave behind it are done (ev_dirs_free, recorded in api.cu), and nothing else of this stream — the previous
        // pulse's thin waves and bin emission, this pulse's pose update — has to be waited for; whatever of the footprint
        // work stays on this stream waits for the directions (join_head)
        const bool side = e->side_dirs && !e->knobs.no_overlap;
        if (side) {
            if (!e->dirs_free_valid) cudaEventRecord(e->ev_dirs_free, st);
            cudaStreamWaitEvent(e->side_dirs, e->ev_dirs_free, 0);
        }
        for (int a = 0; a < 3; a++) p.dirs[a] = e->d_dirs + (size_t)a * e->raster_alloc;
        p.hits = e->d_hits;
        if (tl) cudaEventRecord(tl[0], side ? e->side_dirs : st);
        k_primary_dirs<<<e->num_sms * 32, RTS_DIRS_BLOCK, 0, side ? e->side_dirs : st>>>(p);   // also resets the hit words
        if (tl) cudaEventRecord(tl[1], side ? e->side_dirs : st);
        head_open = side;
        e->dirs_key = key;
        e->dirs_valid = true;
        e->static_valid = false;
        e->launches++;
    }
    // The closest hits among the triangles that never move are the same from pulse to pulse as long as the launch, the
    // scene and the set of moving targets are: they are kept (static pass once), and a pulse only projects the
    // triangles of the moving targets on top of a copy.  Needs the moving-target lists of the partial refit (bvh.cu).
    uint32_t n_moving = 0;
    for (uint32_t k = 0; k < e->n_targets; k++) n_moving += e->moving[k] ? 1u : 0u;
    const bool cacheable = reuse && single_batch && !e->knobs.no_static_hits && (n_moving == 0 || e->partial_ready);
    e->coh_on = false;
    e->split_static = false;
    // From scratch with moving targets: the triangles that never move are projected on side_dirs, behind the direction pass
    // — their records are not touched by the pose update, so nothing of this stream has to be waited for, and the pass
    // (most of the footprint work) runs beside the previous pulse's thin last waves, its bin emission and this pulse's
    // pose update; only the moving targets' triangles are projected here, behind the update.  (A moving triangle's record
    // may be rewritten while the static pass looks at its target id to skip it: the id is the same before and after.)
    // Only when there is something to run beside: an earlier batch of this launch, or a previous pulse still in flight (a
    // host that reads every pulse's bins before it enqueues the next gains nothing, and would pay the second pass's launch
    // latency: 0.07 ms).
    const bool in_flight = p.batch_base > 0 || (e->have_pulse && cudaEventQuery(e->ev[1]) == cudaErrorNotReady);
    const bool split = head_open && !cacheable && n_moving > 0 && e->partial_ready && e->n_dt > 0 && !e->knobs.no_split_raster &&
                       (in_flight || e->knobs.no_split_raster < 0);
    if (tl && !split) cudaEventRecord(tl[2], head_open && !cacheable && n_moving == 0 ? e->side_dirs : st);   // (approximate when the engine's stream still has to wait for the directions)
    if (split) {
        RTS_CUDA(cudaMemsetAsync(e->d_raster_ctl_static, 0, sizeof(RasterCtl), e->side_dirs));
        // the movers' pass needs the directions, the cleared hit words and the cleared control block — not the static pass,
        // beside which it runs (both passes only ever lower hit words; the chunk list is shared out between them)
        cudaEventRecord(e->ev_dirs_done, e->side_dirs);
        const uint32_t mov_items = 1u << 16;
        WaveParams q = p;
        q.raster_ctl = (RasterCtl *)e->d_raster_ctl_static;
        q.raster_skip = e->d_moving;
        q.raster_item_cap = RTS_RASTER_ITEM_CAP - mov_items;
        if (tl) cudaEventRecord(tl[2], e->side_dirs);
        launch_footprints(e, q, p.n_tris, e->side_dirs);
        cudaStreamWaitEvent(st, e->ev_dirs_done, 0);
        p.raster_static = (const RasterCtl *)e->d_raster_ctl_static;
        p.raster_list = e->d_tlist; p.raster_list_count = e->n_dt;
        p.raster_items = (RasterItem *)e->d_raster_items + (RTS_RASTER_ITEM_CAP - mov_items);
        p.raster_item_cap = mov_items;
        launch_footprints(e, p, e->n_dt);
        join_head();          // the shading pass needs both
        e->static_valid = false;
        e->split_static = true;
    } else if (cacheable) {
        join_head();
        const bool valid = e->static_valid && same_launch && e->static_scene_version == e->scene_version &&
                           e->static_moving_version == e->moving_version;
        if (!valid) {
            if (same_launch) RTS_CUDA(cudaMemsetAsync(e->d_hits, 0xff, sizeof(unsigned long long) * p.n_primary, st));
            RTS_CUDA(cudaMemsetAsync(e->d_raster_ctl_static, 0, sizeof(RasterCtl), st));
            WaveParams q = p;
            q.raster_ctl = (RasterCtl *)e->d_raster_ctl_static;
            q.raster_skip = n_moving ? e->d_moving : nullptr;
            launch_footprints(e, q, p.n_tris);
            RTS_CUDA(cudaMemcpyAsync(e->d_hits_static, e->d_hits, sizeof(unsigned long long) * p.n_primary, cudaMemcpyDeviceToDevice, st));
            e->static_valid = true;
            e->static_scene_version = e->scene_version;
            e->static_moving_version = e->moving_version;
        } else {
            RTS_CUDA(cudaMemcpyAsync(e->d_hits, e->d_hits_static, sizeof(unsigned long long) * p.n_primary, cudaMemcpyDeviceToDevice, st));
        }
        p.raster_static = (const RasterCtl *)e->d_raster_ctl_static;
        // kept first-reflection hits (coherent.cuh): same conditions, plus few enough movers to test their boxes one by one
        e->coh_on = n_moving <= 32 && p.dMax >= 2 && !e->knobs.no_kept_reflections;
        if (e->coh_on) {
            const bool w1_ok = valid && e->w1_valid && e->w1_builds == e->builds && e->w1_interp == p.interpolate &&
                               e->w1_dmax == p.dMax && e->w1_rmax == p.rMax;
            e->coh_fill = !w1_ok;
            e->w1_valid = true; e->w1_builds = e->builds; e->w1_interp = p.interpolate; e->w1_dmax = p.dMax; e->w1_rmax = p.rMax;
            p.w1_static = e->d_w1_static;
            p.hits_static = e->d_hits_static;
            p.moving_flags = e->d_moving;
        }
        if (n_moving && e->n_dt) {
            p.raster_list = e->d_tlist; p.raster_list_count = e->n_dt;
            launch_footprints(e, p, e->n_dt);
        }
    } else if (head_open && n_moving == 0) {
        // nothing moves: the whole pass behind the direction pass, beside the previous pulse's last waves (with the control
        // block that is cleared on that stream; the one of this stream stays empty)
        RTS_CUDA(cudaMemsetAsync(e->d_raster_ctl_static, 0, sizeof(RasterCtl), e->side_dirs));
        WaveParams q = p;
        q.raster_ctl = (RasterCtl *)e->d_raster_ctl_static;
        launch_footprints(e, q, p.n_tris, e->side_dirs);
        join_head();
        p.raster_static = (const RasterCtl *)e->d_raster_ctl_static;
        e->static_valid = false;
        e->split_static = true;
    } else {
        join_head();
        if (same_launch) RTS_CUDA(cudaMemsetAsync(e->d_hits, 0xff, sizeof(unsigned long long) * p.n_primary, st));
        e->static_valid = false;
        launch_footprints(e, p, p.n_tris);
    }
    join_head();
    // the resolve pass exists to flag the rays whose first reflection can be answered from kept hits; without that the
    // shading pass looks the leaf position up itself
    p.hits_resolved = e->coh_on ? 1u : 0u;
    if (e->coh_on) { k_raster_resolve<<<e->num_sms * 16, 256, 0, st>>>(p); e->launches++; }
    // Without kept first-reflection hits the shading pass follows each reflection in place (follow.cuh): the second wave's
    // queue round trip (88 bytes written and read back per lit pixel) disappears, and the next launch is a thin one.
    e->followed = !e->coh_on && !e->knobs.no_follow && p.dMax >= 2;
    bvh_join(e);      // the shading pass below is the first kernel of the pulse that may walk the tree
    if (tl) cudaEventRecord(tl[3], st);
    if (e->followed) {
        if (!e->follow_grid) {
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_primary_follow<false>, RTS_WAVE_BLOCK, 0);
            e->follow_grid = e->num_sms * (occ > 0 ? occ : 1);
        }
        trace_wave_grid(e);
        {
            int rc = ensure_coop(e, std::max(e->wave_grid, e->follow_grid));
            if (rc) return rc;
        }
        p.coop_stacks = reinterpret_cast<int *>(e->d_coop_stacks);
        const bool timed = single_batch && e->follow_ev[0];   // this kernel alone, apart from the directions / footprint passes of the wave (rts_get_follow_profile)
        if (timed) cudaEventRecord(e->follow_ev[0], st);
        if (p.rcs_tab || p.ant_rx) k_primary_follow<false, true><<<e->follow_grid, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
        else if (records) k_primary_follow<true><<<e->follow_grid, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
        else k_primary_follow<false><<<e->follow_grid, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
        if (timed) { cudaEventRecord(e->follow_ev[1], st); e->follow_timed = true; }
    } else if (p.rcs_tab || p.ant_rx) k_primary_shade<false, true><<<e->num_sms * RTS_SHADE_MIN_BLOCKS, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
    else if (records) k_primary_shade<true><<<e->num_sms * RTS_SHADE_MIN_BLOCKS, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
    else k_primary_shade<false><<<e->num_sms * RTS_SHADE_MIN_BLOCKS, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
    RTS_CUDA(cudaGetLastError());
    e->launches += 1;
    if (tl) cudaEventRecord(tl[4], st);
    return RTS_OK;
}

// Second wave with kept first-reflection hits: this pulse's boxes of the moving targets, the one-off fill, then the
// kernel that serves the flagged rays; the ordinary wave kernel behind it takes every ray that is not (or no longer) flagged.
int trace_launch_kept(rts_engine *e, WaveParams &p, bool records)
{
    cudaStream_t st = e->stream;
    bvh_join(e);
    MoverIds M;
    memset(&M, 0, sizeof(M));
    for (uint32_t k = 0; k < e->n_targets && M.n < 32; k++)
        if (e->moving[k]) M.id[M.n++] = k;
    if (!e->d_target_box) RTS_CUDA(cudaMalloc(&e->d_target_box, sizeof(unsigned) * 6 * std::max<uint32_t>(1, e->n_targets)));
    if (!e->d_mover_nodes) RTS_CUDA(cudaMalloc(&e->d_mover_nodes, sizeof(BvhNode) * 16));
    p.n_mover_nodes = (M.n + 1) / 2;
    p.mover_nodes = e->d_mover_nodes;
    if (M.n) {
        k_target_box_init<<<1, 192, 0, st>>>(e->d_target_box, M);
        if (e->n_dt) k_target_boxes<<<(e->n_dt + 255) / 256, 256, 0, st>>>(e->d_tlist, e->n_dt, e->d_tri_box, e->d_tri_target, e->d_target_box);
        k_mover_nodes<<<1, 32, 0, st>>>(e->d_target_box, M, e->d_mover_nodes);
        e->launches += 3;
    }
    // to-do list of the ordinary wave kernel: everything the kept kernel does not serve
    if (e->todo_alloc < e->q_capacity) {
        if (e->d_todo) cudaFree(e->d_todo);
        e->d_todo = nullptr; e->todo_alloc = 0;
        RTS_CUDA(cudaMalloc(&e->d_todo, sizeof(uint32_t) * e->q_capacity));
        e->todo_alloc = e->q_capacity;
    }
    p.todo_list = e->d_todo;
    p.todo_count = e->d_counts + 62;
    if (e->coh_fill) {
        RTS_CUDA(cudaMemsetAsync(e->d_w1_static, 0xfe, sizeof(unsigned long long) * p.n_primary, st));
        p.fill_counter = e->d_counts + 63;
        k_wave1_fill<<<e->wave_grid, RTS_WAVE_BLOCK, 0, st>>>(p);
        e->launches++;
    }
    if (p.rcs_tab || p.ant_rx) k_wave1_kept<false, true><<<e->num_sms * RTS_SHADE_MIN_BLOCKS, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
    else if (records) k_wave1_kept<true><<<e->num_sms * RTS_SHADE_MIN_BLOCKS, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
    else k_wave1_kept<false><<<e->num_sms * RTS_SHADE_MIN_BLOCKS, RTS_WAVE_BLOCK, bins_smem_bytes(p), st>>>(p);
    e->launches++;
    RTS_CUDA(cudaGetLastError());
    return RTS_OK;
}

int trace_alloc_queues(rts_engine *e, uint64_t capacity)
{
    if (capacity <= e->q_capacity) return RTS_OK;
    const size_t per = (size_t)RTS_NF * 8 + 8 + 4 + 4;
    for (int k = 0; k < 2; k++) {
        if (e->q_slab[k]) { cudaFree(e->q_slab[k]); e->q_slab[k] = nullptr; }
        cudaError_t err = cudaMalloc(&e->q_slab[k], per * capacity + 256);
        if (err != cudaSuccess) {
            e->q_capacity = 0;
            return rts_fail(RTS_ERR_CUDA, "queue allocation of %zu bytes failed: %s", per * capacity, cudaGetErrorString(err));
        }
        char *base = (char *)e->q_slab[k];
        for (int f = 0; f < RTS_NF; f++) e->q[k].f[f] = (double *)(base + (size_t)f * 8 * capacity);
        e->q[k].key = (unsigned long long *)(base + (size_t)RTS_NF * 8 * capacity);
        e->q[k].ray = (uint32_t *)(base + (size_t)(RTS_NF + 1) * 8 * capacity);
        e->q[k].meta = (uint32_t *)(base + (size_t)(RTS_NF + 1) * 8 * capacity + 4 * capacity);
    }
    e->q_capacity = capacity;
    return RTS_OK;
}
