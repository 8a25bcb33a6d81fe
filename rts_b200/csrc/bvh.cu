// bvh.cu — device-side scene update, Morton LBVH build and bottom-up refit (sm_100a).
//
// Replaces, for the B200 path, the reference's `bound` program (triangle_mesh.cu:204-233) and the
// closed-source OptiX "Bvh" builder it feeds (ray_tracer.cpp:1126-1130), plus the per-pulse host
// rotate/translate of every target (ray_tracer.cpp:156-170, 996-1014).
//
// Pipeline (all on the engine's stream, no host synchronisation):
//   k_transform      base vertices/normals -> world (fp64, op order of the reference, no FMA)
//   k_tri_boxes      per-triangle fp32 AABB rounded outward (cvt.rd / cvt.ru) + scene bounds
//   k_morton         63-bit Morton code of the box centre           } build only
//   cub radix sort   (code, triangle id)                            }
//   k_hierarchy      Karras 2012 binary radix tree                  }
//   k_tri_records    leaf-ordered 80-byte triangle records (128-bit stores)
//   k_fit            bottom-up box union with one atomic counter per internal node
//   k_pack           64-byte traversal nodes holding both child boxes; subtrees of <= 4
//                    triangles are collapsed into one leaf reference
#include "engine.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <vector>
#include <math_constants.h>
#include <cstring>
#include <cmath>

namespace {

__device__ __forceinline__ unsigned f2ord(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// world = (has_rotation ? R*base : base) + t ; normals: rotation only.  Accumulation order of
// matrix_multiply (ray_tracer.cpp:120-137): ((0 + R[i][0]*v0) + R[i][1]*v1) + R[i][2]*v2.
// `list` (optional) restricts the kernel to the elements it names — the vertices of the targets that move.
__global__ void k_transform(const double *__restrict__ base, double *__restrict__ world,
                            const uint32_t *__restrict__ owner, const rts_pose *__restrict__ poses, uint32_t n,
                            int translate, const uint32_t *__restrict__ list)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (list) i = list[i];
    const rts_pose &P = poses[owner[i]];
    double v0 = base[3 * (size_t)i], v1 = base[3 * (size_t)i + 1], v2 = base[3 * (size_t)i + 2];
    double w0 = v0, w1 = v1, w2 = v2;
    if (P.has_rotation) {
        w0 = ((0.0 + P.R[0] * v0) + P.R[1] * v1) + P.R[2] * v2;
        w1 = ((0.0 + P.R[3] * v0) + P.R[4] * v1) + P.R[5] * v2;
        w2 = ((0.0 + P.R[6] * v0) + P.R[7] * v1) + P.R[8] * v2;
    }
    if (translate) {
        w0 += P.t[0]; w1 += P.t[1]; w2 += P.t[2];
    }
    world[3 * (size_t)i] = w0; world[3 * (size_t)i + 1] = w1; world[3 * (size_t)i + 2] = w2;
}

// triangle_mesh.cu:204-233: fp64 min/max, narrowed with directed rounding.
__global__ void k_tri_boxes(const double *__restrict__ world, const uint32_t *__restrict__ tris,
                            const uint32_t *__restrict__ tri_target, const uint32_t *__restrict__ t_vert_off,
                            float *__restrict__ tri_box, unsigned *__restrict__ scene_box, uint32_t n,
                            const uint32_t *__restrict__ list)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    if (i < n) {
        if (list) i = list[i];
        const uint32_t voff = t_vert_off[tri_target[i]];
        double mn[3], mx[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double *v = world + 3 * (size_t)(voff + tris[3 * (size_t)i + k]);
#pragma unroll
            for (int a = 0; a < 3; a++) {
                double x = v[a];
                if (k == 0) { mn[a] = x; mx[a] = x; }
                else { mn[a] = x < mn[a] ? x : mn[a]; mx[a] = x > mx[a] ? x : mx[a]; }
            }
        }
#pragma unroll
        for (int a = 0; a < 3; a++) {
            lo[a] = __double2float_rd(mn[a]);
            hi[a] = __double2float_ru(mx[a]);
            tri_box[6 * (size_t)i + a] = lo[a];
            tri_box[6 * (size_t)i + 3 + a] = hi[a];
        }
    }
    // scene bounds: warp reduce, then one atomic per warp
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float l = lo[a], h = hi[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0 && l <= h) {
            atomicMin(scene_box + a, f2ord(l));
            atomicMax(scene_box + 3 + a, f2ord(h));
        }
    }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v)
{
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton(const float *__restrict__ tri_box, const unsigned *__restrict__ scene_box,
                         unsigned long long *__restrict__ codes, uint32_t *__restrict__ ids, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long code = 0;
    // one scale for all three axes (cubic Morton cells): a flat scene such as a height field must not
    // spend every third split on its thin axis
    float ext = 0.f;
#pragma unroll
    for (int a = 0; a < 3; a++) ext = fmaxf(ext, ord2f(scene_box[3 + a]) - ord2f(scene_box[a]));
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float lo = ord2f(scene_box[a]);
        float c = 0.5f * (tri_box[6 * (size_t)i + a] + tri_box[6 * (size_t)i + 3 + a]);
        float u = ext > 0.f ? (c - lo) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);
        unsigned long long q = (unsigned long long)(u * 2097151.0f);
        code |= expand21(q) << (2 - a);
    }
    codes[i] = code;
    ids[i] = i;
}

__device__ __forceinline__ int delta(const unsigned long long *__restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)(a ^ b));
}

// Karras, "Maximizing parallelism in the construction of BVHs, octrees and k-d trees" (HPG 2012).
__global__ void k_hierarchy(const unsigned long long *__restrict__ keys, int n, int2 *__restrict__ children,
                            int2 *__restrict__ range, int32_t *__restrict__ parent)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? ~gamma : gamma;
    const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    range[i] = make_int2(lo, hi);
    parent[left >= 0 ? left : (n - 1) + ~left] = i;
    parent[right >= 0 ? right : (n - 1) + ~right] = i;
    if (i == 0) parent[0] = -1;
}

#include "ploc.cuh"

__global__ void k_leaf_of_tri(const uint32_t *__restrict__ order, uint32_t *__restrict__ leaf_of_tri, uint32_t n)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) leaf_of_tri[order[p]] = p;
}

// Leaf-ordered triangle records: one thread per leaf position, 5 x 128-bit stores.
__global__ void k_tri_records(const double *__restrict__ world, const uint32_t *__restrict__ tris,
                              const uint32_t *__restrict__ tri_target, const uint32_t *__restrict__ t_vert_off,
                              const uint32_t *__restrict__ order, TriRec *__restrict__ rec, uint32_t n,
                              const uint32_t *__restrict__ list, const uint32_t *__restrict__ leaf_of_tri)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t id;
    if (list) { id = list[p]; p = leaf_of_tri[id]; }   // one thread per moving triangle
    else id = order[p];                                 // one thread per leaf position
    const uint32_t targ = tri_target[id];
    const uint32_t voff = t_vert_off[targ];
    double v[9];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double *s = world + 3 * (size_t)(voff + tris[3 * (size_t)id + k]);
        v[3 * k] = s[0]; v[3 * k + 1] = s[1]; v[3 * k + 2] = s[2];
    }
    double2 *dst = reinterpret_cast<double2 *>(rec + p);
    dst[0] = make_double2(v[0], v[1]);
    dst[1] = make_double2(v[2], v[3]);
    dst[2] = make_double2(v[4], v[5]);
    dst[3] = make_double2(v[6], v[7]);
    unsigned long long idbits = (unsigned long long)id | ((unsigned long long)targ << 32);
    dst[4] = make_double2(v[8], __longlong_as_double((long long)idbits));
}

__device__ __forceinline__ void load_box(const float *p, float b[6])
{
#pragma unroll
    for (int a = 0; a < 6; a++) b[a] = __ldcg(p + a);
}

// Bottom-up fit: the last thread to arrive at a node unions its children and carries on.  Full refit: one thread
// per leaf, two arrivals per node.  Partial refit: one thread per moving triangle (`list`), and a node waits for as
// many arrivals as it has children with a moving triangle below them (`mark`, two bits per node); the boxes of
// untouched children are still valid.
__global__ void k_fit(const int2 *__restrict__ children, const int32_t *__restrict__ parent,
                      const uint32_t *__restrict__ order, const float *tri_box, float *node_box,
                      uint32_t *__restrict__ flags, int n, int count, const uint32_t *__restrict__ list,
                      const uint32_t *__restrict__ leaf_of_tri, const uint32_t *__restrict__ mark)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    if (list) p = (int)leaf_of_tri[list[p]];
    int cur = parent[(n - 1) + p];
    while (cur >= 0) {
        __threadfence();
        const unsigned need = mark ? (unsigned)__popc(mark[cur] & 3u) : 2u;
        if (atomicAdd(flags + cur, 1u) + 1u < need) return;
        const int2 ch = children[cur];
        float a[6], b[6];
        load_box(ch.x >= 0 ? node_box + 6 * (size_t)ch.x : tri_box + 6 * (size_t)order[~ch.x], a);
        load_box(ch.y >= 0 ? node_box + 6 * (size_t)ch.y : tri_box + 6 * (size_t)order[~ch.y], b);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            __stcg(node_box + 6 * (size_t)cur + k, fminf(a[k], b[k]));
            __stcg(node_box + 6 * (size_t)cur + 3 + k, fmaxf(a[k + 3], b[k + 3]));
        }
        cur = parent[cur];
    }
}

// Frame of the quantised nodes: the scene box plus 1/16 of its extent on every side (moving targets may wander that far
// before a box leaves the frame and everything is quantised again), never thinner than 2^-10 of the largest extent.
__global__ void k_qframe(const unsigned *__restrict__ scene_box, QFrame *__restrict__ frame, int only_on_overflow)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (only_on_overflow && !frame->overflow) return;
    float lo[3], hi[3], big = 0.f;
    for (int a = 0; a < 3; a++) {
        lo[a] = ord2f(scene_box[a]); hi[a] = ord2f(scene_box[3 + a]);
        if (!(lo[a] <= hi[a])) { lo[a] = 0.f; hi[a] = 0.f; }
        big = fmaxf(big, hi[a] - lo[a]);
    }
    big = fmaxf(big, 1e-6f);
    for (int a = 0; a < 3; a++) {
        const float e = fmaxf(hi[a] - lo[a], big * 9.765625e-4f);
        frame->lo[a] = lo[a] - 0.0625f * e;
        frame->ext[a] = e * 1.125f;
    }
    // overflow stays set for k_pack_all behind this launch, which clears it
}

__global__ void k_qframe_done(QFrame *frame) { if (threadIdx.x == 0) frame->overflow = 0u; }

// 15-bit plane positions of a box on the frame, rounded outward and widened by one cell (covers the traversal's
// evaluation error, trace.cu: traverse_q); a plane that would leave the frame raises the overflow flag
__device__ __forceinline__ uint32_t q_planes(float lo, float hi, float flo, float ext, uint32_t *overflow)
{
    const double s = 32768.0 / (double)ext;
    const double a = floor(((double)lo - (double)flo) * s) - 1.0, b = ceil(((double)hi - (double)flo) * s) + 1.0;
    if (a < 0.0 || b > 32767.0) *overflow = 1u;
    const uint32_t qa = (uint32_t)fmin(fmax(a, 0.0), 32767.0), qb = (uint32_t)fmin(fmax(b, 0.0), 32767.0);
    return (0x8000u | qa) | ((0x8000u | qb) << 16);
}

__global__ void k_pack(const int2 *__restrict__ children, const int2 *__restrict__ range,
                       const uint32_t *__restrict__ order, const float *__restrict__ tri_box,
                       const float *__restrict__ node_box, BvhNode *__restrict__ nodes, int n, double *sah, int leaf_max,
                       int count, const uint32_t *__restrict__ list, uint32_t *__restrict__ flags,
                       QNode *__restrict__ qnodes, QFrame *__restrict__ frame, int only_on_overflow)
{
    if (only_on_overflow && !*reinterpret_cast<volatile uint32_t *>(&frame->overflow)) return;   // re-quantisation pass: nothing left the frame
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < count;
    if (live && list) { i = (int)list[i]; flags[i] = 0u; }   // partial refit: re-arm the node's arrival counter
    if (!live) i = n;
    // surface-area-heuristic cost of the tree = sum of internal-node box areas (relative to the root's):
    // warp-reduced, one fp64 atomic per warp
    double area = 0.0;
    if (i < n - 1) {
        const float *b = node_box + 6 * (size_t)i;
        const double dx = (double)b[3] - b[0], dy = (double)b[4] - b[1], dz = (double)b[5] - b[2];
        area = dx * dy + dy * dz + dz * dx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(0xffffffffu, area, o);
    if ((threadIdx.x & 31) == 0 && area > 0.0 && !only_on_overflow) atomicAdd(sah, area);
    if (i >= n - 1) return;
    const int2 ch = children[i];
    BvhNode nd;
    QNode qn;
    uint32_t q_over = 0;
    int refs[2];
    float cc[2][3], hh[2][3];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int r = c == 0 ? ch.x : ch.y;
        const float *src;
        if (r < 0) {
            refs[c] = ~(((~r) << 3) | 0);
            src = tri_box + 6 * (size_t)order[~r];
        } else {
            const int2 rg = range[r];
            const int cnt = rg.y - rg.x + 1;
            refs[c] = cnt <= leaf_max ? ~((rg.x << 3) | (cnt - 1)) : r;
            src = node_box + 6 * (size_t)r;
        }
#pragma unroll
        for (int a = 0; a < 3; a++) {
            // centre / half-extent with [c-h, c+h] >= [lo, hi]: the differences are evaluated in fp64 (relative
            // error 2^-53), rounded up to fp32 and bumped one more ulp, so containment holds exactly
            const float lo = src[a], hi = src[3 + a];
            const float c0 = 0.5f * lo + 0.5f * hi;
            const double d = fmax((double)hi - (double)c0, (double)c0 - (double)lo);
            cc[c][a] = c0;
            hh[c][a] = nextafterf(__double2float_ru(d), CUDART_INF_F);
            if (qnodes) qn.w[3 * c + a] = q_planes(lo, hi, frame->lo[a], frame->ext[a], &q_over);
        }
    }
    if (qnodes) {      // the quantised copy of the node: only when a kernel that walks it is enabled (want_qnodes)
        qn.ref0 = refs[0]; qn.ref1 = refs[1];
        uint4 *qd = reinterpret_cast<uint4 *>(qnodes + i);
        const uint4 *qs = reinterpret_cast<const uint4 *>(&qn);
        qd[0] = qs[0]; qd[1] = qs[1];
        if (q_over && !only_on_overflow) atomicOr(&frame->overflow, 1u);
    }
    if (only_on_overflow) return;          // the fp32 nodes and the SAH sum are already in place
    nd.c0x = cc[0][0]; nd.c0y = cc[0][1]; nd.h0x = hh[0][0]; nd.h0y = hh[0][1];
    nd.c1x = cc[1][0]; nd.c1y = cc[1][1]; nd.h1x = hh[1][0]; nd.h1y = hh[1][1];
    nd.c0z = cc[0][2]; nd.c1z = cc[1][2]; nd.h0z = hh[0][2]; nd.h1z = hh[1][2];
    nd.ref0 = refs[0]; nd.ref1 = refs[1]; nd.pad[0] = nd.pad[1] = 0;
    float4 *dst = reinterpret_cast<float4 *>(nodes + i);
    const float4 *s4 = reinterpret_cast<const float4 *>(&nd);
    dst[0] = s4[0]; dst[1] = s4[1]; dst[2] = s4[2]; dst[3] = s4[3];
}

// The quantised node copy (QNode) is maintained only when something walks it: the split form of the later waves
// (k_traverse, option no_split = 0) or a tuning build with RTS_QNODES = 1.
static bool want_qnodes(const rts_engine *e) { return RTS_QNODES != 0 || !e->knobs.no_split; }

// ---- partial refit set-up (run when the set of moving targets changes) ----
// Elements owned by a moving target, appended in arbitrary order (one atomic per warp).
__global__ void k_select_owned(const uint32_t *__restrict__ owner, const uint32_t *__restrict__ moving, uint32_t n,
                               uint32_t *__restrict__ out, uint32_t *__restrict__ count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool take = i < n && moving[owner[i]] != 0u;
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (!m) return;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (take) out[base + __popc(m & ((1u << lane) - 1u))] = i;
}
// mark[node] bit 0 / bit 1: child 0 / child 1 has a moving triangle below it.
__global__ void k_mark_paths(const uint32_t *__restrict__ list, uint32_t count, const uint32_t *__restrict__ leaf_of_tri,
                             const int2 *__restrict__ children, const int32_t *__restrict__ parent, int n,
                             uint32_t *__restrict__ mark)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int p = (int)leaf_of_tri[list[k]];
    int from = ~p;
    int cur = parent[(n - 1) + p];
    while (cur >= 0) {
        const unsigned bit = children[cur].x == from ? 1u : 2u;
        if (atomicOr(mark + cur, bit) & bit) return;   // whoever set it first carries on upward
        from = cur;
        cur = parent[cur];
    }
}
__global__ void k_select_marked(const uint32_t *__restrict__ mark, uint32_t n, uint32_t *__restrict__ out,
                                uint32_t *__restrict__ count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool take = i < n && mark[i] != 0u;
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (!m) return;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (take) out[base + __popc(m & ((1u << lane) - 1u))] = i;
}
// Scene bounds of the triangles that never move and SAH area of the nodes no moving triangle touches.
__global__ void k_static_box(const float *__restrict__ tri_box, const uint32_t *__restrict__ tri_target,
                             const uint32_t *__restrict__ moving, uint32_t n, unsigned *__restrict__ box)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    if (i < n && moving[tri_target[i]] == 0u) {
#pragma unroll
        for (int a = 0; a < 3; a++) { lo[a] = tri_box[6 * (size_t)i + a]; hi[a] = tri_box[6 * (size_t)i + 3 + a]; }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float l = lo[a], h = hi[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0 && l <= h) {
            atomicMin(box + a, f2ord(l));
            atomicMax(box + 3 + a, f2ord(h));
        }
    }
}
__global__ void k_static_sah(const float *__restrict__ node_box, const uint32_t *__restrict__ mark, int n, double *sah)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double area = 0.0;
    if (i < n - 1 && mark[i] == 0u) {
        const float *b = node_box + 6 * (size_t)i;
        const double dx = (double)b[3] - b[0], dy = (double)b[4] - b[1], dz = (double)b[5] - b[2];
        area = dx * dy + dy * dz + dz * dx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(0xffffffffu, area, o);
    if ((threadIdx.x & 31) == 0 && area > 0.0) atomicAdd(sah, area);
}
// max |coordinate| of the scene box per axis, for the slab-test error bound of the traversal (trace.cu)
__global__ void k_scene_abs(const unsigned *__restrict__ box, float *__restrict__ out, uint32_t n_tris)
{
    const int a = threadIdx.x;
    if (a >= 3) return;
    const float m = fmaxf(fabsf(ord2f(box[a])), fabsf(ord2f(box[3 + a])));
    out[a] = (n_tris && m < 3.0e38f) ? m : 0.f;
}

// Invariant check: every triangle box inside all of its ancestors' boxes.
__global__ void k_check(const int32_t *__restrict__ parent, const uint32_t *__restrict__ order,
                        const float *__restrict__ tri_box, const float *__restrict__ node_box, int n,
                        unsigned long long *violations)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float b[6];
    for (int a = 0; a < 6; a++) b[a] = tri_box[6 * (size_t)order[p] + a];
    int cur = parent[(n - 1) + p];
    unsigned bad = 0;
    while (cur >= 0) {
        const float *nb = node_box + 6 * (size_t)cur;
        for (int a = 0; a < 3; a++) bad += (b[a] < nb[a]) + (b[a + 3] > nb[a + 3]);
        cur = parent[cur];
    }
    if (bad) atomicAdd(violations, (unsigned long long)bad);
}

inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

template <class T> int dalloc(T **p, size_t count)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (!count) count = 1;
    cudaError_t err = cudaMalloc((void **)p, count * sizeof(T));
    if (err != cudaSuccess) return rts_fail(RTS_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(err));
    return RTS_OK;
}

} // namespace

void bvh_free(rts_engine *e)
{
    void **ptrs[] = {(void **)&e->d_morton, (void **)&e->d_morton_sorted, (void **)&e->d_order_in, (void **)&e->d_order,
                     (void **)&e->d_leaf_of_tri, (void **)&e->d_tri_box, (void **)&e->d_node_box, (void **)&e->d_scene_box,
                     (void **)&e->d_parent, (void **)&e->d_children, (void **)&e->d_range, (void **)&e->d_fit_flags,
                     (void **)&e->d_nodes, (void **)&e->d_trirec, (void **)&e->d_cub_temp, (void **)&e->d_violations, (void **)&e->d_sah,
                     (void **)&e->d_moving, (void **)&e->d_vlist, (void **)&e->d_nlist, (void **)&e->d_tlist, (void **)&e->d_nodelist,
                     (void **)&e->d_list_counts, (void **)&e->d_mark, (void **)&e->d_static_box, (void **)&e->d_sah_static,
                     (void **)&e->d_scene_abs, (void **)&e->d_qnodes, (void **)&e->d_qframe};
    for (void **p : ptrs) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    e->partial_ready = false;
    e->sah_pending = false;
}

int bvh_alloc(rts_engine *e)
{
    const size_t T = e->n_tris;
    int rc;
    if ((rc = dalloc(&e->d_morton, T))) return rc;
    if ((rc = dalloc(&e->d_morton_sorted, T))) return rc;
    if ((rc = dalloc(&e->d_order_in, T))) return rc;
    if ((rc = dalloc(&e->d_order, T))) return rc;
    if ((rc = dalloc(&e->d_leaf_of_tri, T))) return rc;
    if ((rc = dalloc(&e->d_tri_box, 6 * T))) return rc;
    if ((rc = dalloc(&e->d_node_box, 6 * T))) return rc;
    if ((rc = dalloc(&e->d_scene_box, (size_t)6))) return rc;
    if ((rc = dalloc(&e->d_parent, 2 * T))) return rc;
    if ((rc = dalloc(&e->d_children, T))) return rc;
    if ((rc = dalloc(&e->d_range, T))) return rc;
    if ((rc = dalloc(&e->d_fit_flags, T))) return rc;
    if ((rc = dalloc(&e->d_nodes, T))) return rc;
    if ((rc = dalloc(&e->d_qnodes, T))) return rc;
    if ((rc = dalloc(&e->d_qframe, 1))) return rc;
    RTS_CUDA(cudaMemset(e->d_qframe, 0, sizeof(QFrame)));
    if ((rc = dalloc(&e->d_trirec, T))) return rc;
    if ((rc = dalloc(&e->d_violations, (size_t)1))) return rc;
    if ((rc = dalloc(&e->d_sah, (size_t)1))) return rc;
    // partial refit (only the targets that move)
    if ((rc = dalloc(&e->d_moving, (size_t)e->n_targets))) return rc;
    RTS_CUDA(cudaMemsetAsync(e->d_moving, 0, sizeof(uint32_t) * std::max<size_t>(1, e->n_targets), e->stream));
    if ((rc = dalloc(&e->d_vlist, (size_t)e->n_verts))) return rc;
    if ((rc = dalloc(&e->d_nlist, (size_t)e->n_normals))) return rc;
    if ((rc = dalloc(&e->d_tlist, T))) return rc;
    if ((rc = dalloc(&e->d_nodelist, T))) return rc;
    if ((rc = dalloc(&e->d_list_counts, (size_t)4))) return rc;
    if ((rc = dalloc(&e->d_mark, T))) return rc;
    if ((rc = dalloc(&e->d_static_box, (size_t)6))) return rc;
    if ((rc = dalloc(&e->d_sah_static, (size_t)1))) return rc;
    if ((rc = dalloc(&e->d_scene_abs, (size_t)4))) return rc;
    RTS_CUDA(cudaMemsetAsync(e->d_scene_abs, 0, sizeof(float) * 4, e->stream));
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, e->d_morton, e->d_morton_sorted, e->d_order_in, e->d_order, (int)T, 0,
                                    63, e->stream);
    e->cub_temp_bytes = bytes;
    if (e->d_cub_temp) { cudaFree(e->d_cub_temp); e->d_cub_temp = nullptr; }
    RTS_CUDA(cudaMalloc(&e->d_cub_temp, bytes ? bytes : 16));
    return RTS_OK;
}

static const unsigned kEmptyBox[6] = {0xff800000u, 0xff800000u, 0xff800000u, 0x007fffffu, 0x007fffffu, 0x007fffffu}; // lo = +inf, hi = -inf (ordered encoding)

// Every vertex, normal and leaf box from the current poses (all targets).
int bvh_update_world(rts_engine *e)
{
    const unsigned bs = 256;
    if (e->n_verts)
        { k_transform<<<blocks_for(e->n_verts, bs), bs, 0, e->stream>>>(e->d_base_verts, e->d_world_verts, e->d_vert_target,
                                                                     e->d_poses, e->n_verts, 1, nullptr); e->launches++; }
    if (e->n_normals)
        { k_transform<<<blocks_for(e->n_normals, bs), bs, 0, e->stream>>>(e->d_base_normals, e->d_world_normals,
                                                                       e->d_norm_target, e->d_poses, e->n_normals, 0, nullptr); e->launches++; }
    RTS_CUDA(cudaMemcpyAsync(e->d_scene_box, kEmptyBox, sizeof(kEmptyBox), cudaMemcpyHostToDevice, e->stream));
    if (e->n_tris)
        { k_tri_boxes<<<blocks_for(e->n_tris, bs), bs, 0, e->stream>>>(e->d_world_verts, e->d_tris, e->d_tri_target,
                                                                    e->d_t_vert_off, e->d_tri_box,
                                                                    (unsigned *)e->d_scene_box, e->n_tris, nullptr); e->launches++; }
    { k_scene_abs<<<1, 32, 0, e->stream>>>((const unsigned *)e->d_scene_box, e->d_scene_abs, e->n_tris); e->launches++; }
    RTS_CUDA(cudaGetLastError());
    return RTS_OK;
}

static int fit_and_pack(rts_engine *e)
{
    const unsigned bs = 256;
    const int n = (int)e->n_tris;
    if (n == 0) { e->root_ref = 0; return RTS_OK; }
    { k_tri_records<<<blocks_for(n, bs), bs, 0, e->stream>>>(e->d_world_verts, e->d_tris, e->d_tri_target, e->d_t_vert_off,
                                                          e->d_order, e->d_trirec, n, nullptr, nullptr); e->launches++; }
    if (n >= 2) {
        RTS_CUDA(cudaMemsetAsync(e->d_fit_flags, 0, sizeof(uint32_t) * (size_t)n, e->stream));
        { k_fit<<<blocks_for(n, bs), bs, 0, e->stream>>>(e->d_children, e->d_parent, e->d_order, e->d_tri_box, e->d_node_box,
                                                      e->d_fit_flags, n, n, nullptr, nullptr, nullptr); e->launches++; }
        RTS_CUDA(cudaMemsetAsync(e->d_fit_flags, 0, sizeof(uint32_t) * (size_t)n, e->stream));   // armed for partial refits
        RTS_CUDA(cudaMemsetAsync(e->d_sah, 0, sizeof(double), e->stream));
        if (want_qnodes(e)) {
            RTS_CUDA(cudaMemsetAsync(&e->d_qframe->overflow, 0, sizeof(uint32_t), e->stream));
            { k_qframe<<<1, 32, 0, e->stream>>>((const unsigned *)e->d_scene_box, e->d_qframe, 0); e->launches++; }
        }
        { k_pack<<<blocks_for(n - 1, bs), bs, 0, e->stream>>>(e->d_children, e->d_range, e->d_order, e->d_tri_box,
                                                           e->d_node_box, e->d_nodes, n, e->d_sah, e->leaf_max, n - 1, nullptr,
                                                           nullptr, want_qnodes(e) ? e->d_qnodes : nullptr, e->d_qframe, 0); e->launches++; }
    }
    RTS_CUDA(cudaGetLastError());
    e->root_ref = n <= e->leaf_max ? ~((0 << 3) | (n - 1)) : 0;
    return RTS_OK;
}

// Lists and constants of the partial refit for the current set of moving targets (e->moving).  Runs when that set
// changes or the topology was rebuilt; synchronises (it reads four counts back).
static int prepare_partial(rts_engine *e)
{
    const unsigned bs = 256;
    const int n = (int)e->n_tris;
    std::vector<uint32_t> mv(e->n_targets);
    for (uint32_t k = 0; k < e->n_targets; k++) mv[k] = e->moving[k] ? 1u : 0u;
    RTS_CUDA(cudaMemcpyAsync(e->d_moving, mv.data(), sizeof(uint32_t) * e->n_targets, cudaMemcpyHostToDevice, e->stream));
    RTS_CUDA(cudaMemsetAsync(e->d_list_counts, 0, sizeof(uint32_t) * 4, e->stream));
    RTS_CUDA(cudaMemsetAsync(e->d_mark, 0, sizeof(uint32_t) * (size_t)n, e->stream));
    RTS_CUDA(cudaMemsetAsync(e->d_sah_static, 0, sizeof(double), e->stream));
    RTS_CUDA(cudaMemcpyAsync(e->d_static_box, kEmptyBox, sizeof(kEmptyBox), cudaMemcpyHostToDevice, e->stream));
    if (e->n_verts) { k_select_owned<<<blocks_for(e->n_verts, bs), bs, 0, e->stream>>>(e->d_vert_target, e->d_moving, e->n_verts, e->d_vlist, e->d_list_counts + 0); e->launches++; }
    if (e->n_normals) { k_select_owned<<<blocks_for(e->n_normals, bs), bs, 0, e->stream>>>(e->d_norm_target, e->d_moving, e->n_normals, e->d_nlist, e->d_list_counts + 1); e->launches++; }
    { k_select_owned<<<blocks_for(n, bs), bs, 0, e->stream>>>(e->d_tri_target, e->d_moving, n, e->d_tlist, e->d_list_counts + 2); e->launches++; }
    uint32_t counts[4] = {0, 0, 0, 0};
    RTS_CUDA(cudaMemcpyAsync(counts, e->d_list_counts, sizeof(uint32_t) * 3, cudaMemcpyDeviceToHost, e->stream));
    RTS_CUDA(cudaStreamSynchronize(e->stream));   // also makes mv[] safe to drop
    e->n_dv = counts[0]; e->n_dn = counts[1]; e->n_dt = counts[2];
    if (e->n_dt) { k_mark_paths<<<blocks_for(e->n_dt, bs), bs, 0, e->stream>>>(e->d_tlist, e->n_dt, e->d_leaf_of_tri, e->d_children, e->d_parent, n, e->d_mark); e->launches++; }
    { k_select_marked<<<blocks_for(n - 1, bs), bs, 0, e->stream>>>(e->d_mark, n - 1, e->d_nodelist, e->d_list_counts + 3); e->launches++; }
    { k_static_box<<<blocks_for(n, bs), bs, 0, e->stream>>>(e->d_tri_box, e->d_tri_target, e->d_moving, n, (unsigned *)e->d_static_box); e->launches++; }
    { k_static_sah<<<blocks_for(n - 1, bs), bs, 0, e->stream>>>(e->d_node_box, e->d_mark, n, e->d_sah_static); e->launches++; }
    RTS_CUDA(cudaGetLastError());
    RTS_CUDA(cudaMemcpyAsync(counts + 3, e->d_list_counts + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    e->n_dnode = counts[3];
    e->partial_ready = true;
    return RTS_OK;
}

static void collect_sah(rts_engine *e, bool wait);

// Only the moving targets: their vertices, normals, leaf boxes and triangle records, then the tree paths above them.
static int partial_update(rts_engine *e)
{
    const unsigned bs = 256;
    const int n = (int)e->n_tris;
    if (e->n_dv) { k_transform<<<blocks_for(e->n_dv, bs), bs, 0, e->stream>>>(e->d_base_verts, e->d_world_verts, e->d_vert_target, e->d_poses, e->n_dv, 1, e->d_vlist); e->launches++; }
    if (e->n_dn) { k_transform<<<blocks_for(e->n_dn, bs), bs, 0, e->stream>>>(e->d_base_normals, e->d_world_normals, e->d_norm_target, e->d_poses, e->n_dn, 0, e->d_nlist); e->launches++; }
    RTS_CUDA(cudaMemcpyAsync(e->d_scene_box, e->d_static_box, sizeof(unsigned) * 6, cudaMemcpyDeviceToDevice, e->stream));
    if (e->n_dt) {
        { k_tri_boxes<<<blocks_for(e->n_dt, bs), bs, 0, e->stream>>>(e->d_world_verts, e->d_tris, e->d_tri_target, e->d_t_vert_off, e->d_tri_box, (unsigned *)e->d_scene_box, e->n_dt, e->d_tlist); e->launches++; }
        { k_tri_records<<<blocks_for(e->n_dt, bs), bs, 0, e->stream>>>(e->d_world_verts, e->d_tris, e->d_tri_target, e->d_t_vert_off, e->d_order, e->d_trirec, e->n_dt, e->d_tlist, e->d_leaf_of_tri); e->launches++; }
    }
    // Vertices, normals, leaf boxes and triangle records are what the projected primary wave reads; everything from here
    // on — node boxes, packed nodes, scene_abs, the SAH cost — is read by traversal only.  It runs on side_bvh, beside the
    // footprint kernels of the pulse that follows; the first kernel that walks the tree waits for it (bvh_join).  The fork
    // is behind this stream's earlier work, so the previous pulse's last waves are done with the nodes.
    const cudaStream_t main_stream = e->stream;
    const bool side = e->side_bvh && !e->knobs.no_overlap;
    struct Restore { rts_engine *e; cudaStream_t s; ~Restore() { e->stream = s; } } restore{e, main_stream};   // also on the error returns
    if (side) {
        cudaEventRecord(e->ev_bvh_fork, main_stream);
        cudaStreamWaitEvent(e->side_bvh, e->ev_bvh_fork, 0);
        e->stream = e->side_bvh;
    }
    RTS_CUDA(cudaMemcpyAsync(e->d_sah, e->d_sah_static, sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    if (e->n_dt) {
        { k_fit<<<blocks_for(e->n_dt, bs), bs, 0, e->stream>>>(e->d_children, e->d_parent, e->d_order, e->d_tri_box, e->d_node_box, e->d_fit_flags, n, (int)e->n_dt, e->d_tlist, e->d_leaf_of_tri, e->d_mark); e->launches++; }
    }
    if (e->n_dnode) {
        const bool wq = want_qnodes(e);
        k_pack<<<blocks_for(e->n_dnode, bs), bs, 0, e->stream>>>(e->d_children, e->d_range, e->d_order, e->d_tri_box, e->d_node_box, e->d_nodes, n, e->d_sah, e->leaf_max, (int)e->n_dnode, e->d_nodelist, e->d_fit_flags, wq ? e->d_qnodes : nullptr, e->d_qframe, 0);
        e->launches++;
        if (wq) {
        // a moving target has left the frame of the quantised nodes (rare: the frame has a margin of 1/16 of the scene):
        // new frame from this pulse's scene box, every node quantised again.  Both launches return at once otherwise —
        // decided on the device, no host synchronisation.
        k_qframe<<<1, 32, 0, e->stream>>>((const unsigned *)e->d_scene_box, e->d_qframe, 1);
        k_pack<<<blocks_for(n - 1, bs), bs, 0, e->stream>>>(e->d_children, e->d_range, e->d_order, e->d_tri_box, e->d_node_box, e->d_nodes, n, e->d_sah, e->leaf_max, n - 1, nullptr, nullptr, e->d_qnodes, e->d_qframe, 1);
        k_qframe_done<<<1, 32, 0, e->stream>>>(e->d_qframe);
        e->launches += 3;
        }
    }
    { k_scene_abs<<<1, 32, 0, e->stream>>>((const unsigned *)e->d_scene_box, e->d_scene_abs, e->n_tris); e->launches++; }
    if (side) {
        // the SAH cost's read-back belongs behind k_pack: on the side stream too
        if (n >= 2) {
            collect_sah(e, true);   // one read-back slot: the previous one has long arrived
            cudaMemcpyAsync(&e->h_rb->sah, e->d_sah, sizeof(double), cudaMemcpyDeviceToHost, e->stream);
            cudaEventRecord(e->sah_ev, e->stream);
            e->sah_pending = true;
        }
        cudaEventRecord(e->ev_bvh_done, e->side_bvh);
        e->bvh_join_pending = true;
        e->stream = main_stream;
    }
    RTS_CUDA(cudaGetLastError());
    return RTS_OK;
}

// The engine's stream waits for the refit that may still be running on side_bvh.  Called ahead of everything that reads or
// writes the tree: the wave kernels (trace.cu), the next update, a rebuild, the diagnostics.
void bvh_join(rts_engine *e)
{
    if (!e->bvh_join_pending) return;
    cudaStreamWaitEvent(e->stream, e->ev_bvh_done, 0);
    e->bvh_join_pending = false;
}

// Scene box of the current geometry into bvh_info (synchronises; diagnostics only — the kernels read d_scene_abs).
int bvh_read_scene_box(rts_engine *e)
{
    unsigned sb[6];
    RTS_CUDA(cudaMemcpyAsync(sb, e->d_scene_box, sizeof(sb), cudaMemcpyDeviceToHost, e->stream));
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    rts_bvh_info &bi = e->bvh_info;
    for (int a = 0; a < 3; a++) {
        const unsigned lo = sb[a], hi = sb[3 + a];
        const uint32_t l = (lo & 0x80000000u) ? (lo & 0x7fffffffu) : ~lo, h = (hi & 0x80000000u) ? (hi & 0x7fffffffu) : ~hi;
        memcpy(&bi.scene_lo[a], &l, 4);
        memcpy(&bi.scene_hi[a], &h, 4);
    }
    return RTS_OK;
}

// Topology by PLOC (ploc.cuh) over the Morton-sorted triangles in d_order; leaves d_children / d_parent / d_range /
// d_order in the layout of k_hierarchy.  Build-time only: temporaries are allocated and freed here, and each round
// reads two counts back.
static int build_ploc(rts_engine *e)
{
    const int n = (int)e->n_tris;
    const unsigned bs = 256;
    cudaStream_t st = e->stream;
    int *ref[2] = {nullptr, nullptr}, *nn = nullptr, *keep = nullptr, *merged = nullptr, *keep_scan = nullptr, *merged_scan = nullptr;
    int *count = nullptr, *pos_of0 = nullptr;
    int32_t *parent_node = nullptr, *parent_leaf0 = nullptr;
    uint32_t *order_tmp = nullptr;
    PlocBox *box[2] = {nullptr, nullptr};
    void *scan_tmp = nullptr;
    size_t scan_bytes = 0;
    int rc = RTS_OK;
    int m = n, next_id = n - 2, cur = 0, rounds = 0;
#define PLOC_CUDA(call)                                                                                         \
    do {                                                                                                        \
        cudaError_t _e = (call);                                                                                \
        if (_e != cudaSuccess) {                                                                                \
            rc = rts_fail(RTS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            goto done;                                                                                          \
        }                                                                                                       \
    } while (0)
    for (int k = 0; k < 2; k++) {
        PLOC_CUDA(cudaMalloc(&ref[k], sizeof(int) * (size_t)n));
        PLOC_CUDA(cudaMalloc(&box[k], sizeof(PlocBox) * (size_t)n));
    }
    PLOC_CUDA(cudaMalloc(&nn, sizeof(int) * (size_t)n));
    PLOC_CUDA(cudaMalloc(&keep, sizeof(int) * (size_t)n));
    PLOC_CUDA(cudaMalloc(&merged, sizeof(int) * (size_t)n));
    PLOC_CUDA(cudaMalloc(&keep_scan, sizeof(int) * (size_t)n));
    PLOC_CUDA(cudaMalloc(&merged_scan, sizeof(int) * (size_t)n));
    PLOC_CUDA(cudaMalloc(&count, sizeof(int) * (size_t)n));
    PLOC_CUDA(cudaMalloc(&pos_of0, sizeof(int) * (size_t)n));
    PLOC_CUDA(cudaMalloc(&parent_node, sizeof(int32_t) * (size_t)n));
    PLOC_CUDA(cudaMalloc(&parent_leaf0, sizeof(int32_t) * (size_t)n));
    PLOC_CUDA(cudaMalloc(&order_tmp, sizeof(uint32_t) * (size_t)n));
    PLOC_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, keep, keep_scan, n, st));
    PLOC_CUDA(cudaMalloc(&scan_tmp, std::max<size_t>(16, scan_bytes)));
    PLOC_CUDA(cudaMemsetAsync(parent_node, 0xff, sizeof(int32_t) * (size_t)n, st));
    PLOC_CUDA(cudaMemsetAsync(parent_leaf0, 0xff, sizeof(int32_t) * (size_t)n, st));
    k_ploc_init<<<blocks_for(n, bs), bs, 0, st>>>(e->d_order, e->d_tri_box, n, ref[0], box[0]);
    e->launches++;
    while (m > 1) {
        k_ploc_nn<<<blocks_for(m, bs), bs, 0, st>>>(box[cur], m, nn);
        k_ploc_flags<<<blocks_for(m, bs), bs, 0, st>>>(nn, m, keep, merged);
        PLOC_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, keep, keep_scan, m, st));
        PLOC_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, merged, merged_scan, m, st));
        k_ploc_merge<<<blocks_for(m, bs), bs, 0, st>>>(nn, keep_scan, merged_scan, keep, merged, m, next_id, ref[cur], box[cur],
                                                       ref[cur ^ 1], box[cur ^ 1], e->d_children, parent_node, parent_leaf0);
        e->launches += 5;
        int tail[4];
        PLOC_CUDA(cudaMemcpyAsync(tail + 0, keep_scan + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
        PLOC_CUDA(cudaMemcpyAsync(tail + 1, keep + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
        PLOC_CUDA(cudaMemcpyAsync(tail + 2, merged_scan + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
        PLOC_CUDA(cudaMemcpyAsync(tail + 3, merged + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
        PLOC_CUDA(cudaStreamSynchronize(st));
        const int m2 = tail[0] + tail[1], k = tail[2] + tail[3];
        if (k <= 0 || m2 != m - k) { rc = rts_fail(RTS_ERR_STATE, "PLOC round %d made no progress (m=%d, merges=%d)", rounds, m, k); goto done; }
        next_id -= k;
        m = m2;
        cur ^= 1;
        rounds++;
    }
    if (next_id != -1) { rc = rts_fail(RTS_ERR_STATE, "PLOC produced %d internal nodes for %d leaves", n - 2 - next_id, n); goto done; }
    PLOC_CUDA(cudaMemsetAsync(e->d_fit_flags, 0, sizeof(uint32_t) * (size_t)n, st));
    k_ploc_counts<<<blocks_for(n, bs), bs, 0, st>>>(e->d_children, parent_node, parent_leaf0, n, e->d_fit_flags, count);
    k_ploc_positions<<<blocks_for(n, bs), bs, 0, st>>>(e->d_children, parent_node, parent_leaf0, count, n, e->d_order, order_tmp, pos_of0);
    k_ploc_ranges<<<blocks_for(n, bs), bs, 0, st>>>(e->d_children, pos_of0, count, n, e->d_range);
    k_ploc_finish<<<blocks_for(n, bs), bs, 0, st>>>(e->d_children, parent_node, parent_leaf0, pos_of0, n, e->d_parent);
    e->launches += 4;
    PLOC_CUDA(cudaGetLastError());
    PLOC_CUDA(cudaMemcpyAsync(e->d_order, order_tmp, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    PLOC_CUDA(cudaStreamSynchronize(st));
done:
#undef PLOC_CUDA
    for (int k = 0; k < 2; k++) { cudaFree(ref[k]); cudaFree(box[k]); }
    cudaFree(nn); cudaFree(keep); cudaFree(merged); cudaFree(keep_scan); cudaFree(merged_scan); cudaFree(count); cudaFree(pos_of0);
    cudaFree(parent_node); cudaFree(parent_leaf0); cudaFree(order_tmp); cudaFree(scan_tmp);
    return rc;
}

// One topology build at the current world geometry + fit + pack; returns the tree's SAH cost (sum of the
// internal-node box areas).  Synchronises.
static int build_once(rts_engine *e, bool ploc, double *sah_out)
{
    const unsigned bs = 256;
    const int n = (int)e->n_tris;
    int rc;
    if (n > 0) {
        { k_morton<<<blocks_for(n, bs), bs, 0, e->stream>>>(e->d_tri_box, (const unsigned *)e->d_scene_box, e->d_morton,
                                                         e->d_order_in, n); e->launches++; }
        size_t bytes = e->cub_temp_bytes;
        RTS_CUDA(cub::DeviceRadixSort::SortPairs(e->d_cub_temp, bytes, e->d_morton, e->d_morton_sorted, e->d_order_in,
                                                 e->d_order, n, 0, 63, e->stream));
        if (n >= 2) {
            if (ploc) {
                if ((rc = build_ploc(e))) return rc;
            } else {
                k_hierarchy<<<blocks_for(n - 1, bs), bs, 0, e->stream>>>(e->d_morton_sorted, n, e->d_children, e->d_range, e->d_parent);
                e->launches++;
            }
        }
        { k_leaf_of_tri<<<blocks_for(n, bs), bs, 0, e->stream>>>(e->d_order, e->d_leaf_of_tri, n); e->launches++; }
        RTS_CUDA(cudaGetLastError());
    }
    if ((rc = fit_and_pack(e))) return rc;
    double sah = 0;
    if (n >= 2) RTS_CUDA(cudaMemcpyAsync(&sah, e->d_sah, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    *sah_out = sah;
    return RTS_OK;
}

// Full build at the current poses.  Two topology builders feed the same arrays: the Morton radix tree (k_hierarchy,
// < 1 ms per million triangles) and PLOC (ploc.cuh, tens of ms).  Neither wins everywhere — the radix tree is the
// better one on regular height fields, PLOC by far on irregular meshes with mixed triangle sizes (ship on a sea
// plane: 7x lower SAH cost) — so the first build of a scene makes both and keeps the one with the lower SAH cost
// (RTS_BVH=lbvh|ploc forces one); later rebuilds (SAH drift, rts_scene_rebuild) reuse that choice.
int bvh_build(rts_engine *e)
{
    const int n = (int)e->n_tris;
    bvh_join(e);
    e->partial_ready = false;
    e->sah_pending = false;
    cudaEventRecord(e->ev[4], e->stream);
    int rc = bvh_update_world(e);
    if (rc) return rc;
    double sah = 0;
    if (e->builder == 0 && n >= 64) {            // undecided: try both
        double sah_l = 0, sah_p = 0;
        if ((rc = build_once(e, false, &sah_l))) return rc;
        if ((rc = build_once(e, true, &sah_p))) return rc;
        if (sah_p < 0.97 * sah_l) { e->builder = 2; sah = sah_p; }
        else {
            e->builder = 1;
            if ((rc = build_once(e, false, &sah))) return rc;
        }
    } else {
        if ((rc = build_once(e, e->builder == 2, &sah))) return rc;
    }
    cudaEventRecord(e->ev[5], e->stream);
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev[4], e->ev[5]);
    if ((rc = bvh_read_scene_box(e))) return rc;
    rts_bvh_info &bi = e->bvh_info;
    bi.n_tris = e->n_tris;
    bi.n_nodes = n >= 2 ? n - 1 : 0;
    bi.root_is_leaf = e->root_ref < 0;
    bi.max_leaf = e->leaf_max;
    bi.ms_build = ms;
    bi.sah_cost = sah;
    bi.builder = (uint32_t)e->builder;
    e->sah_at_build = sah;
    e->builds++;
    e->refits_since_build = 0;
    return RTS_OK;
}

// Collect the SAH cost the previous refit left in the pinned read-back block, if it has arrived.
static void collect_sah(rts_engine *e, bool wait)
{
    if (!e->sah_pending) return;
    if (wait) cudaEventSynchronize(e->sah_ev);
    else if (cudaEventQuery(e->sah_ev) != cudaSuccess) return;
    e->bvh_info.sah_cost = e->h_rb->sah;
    e->sah_pending = false;
}

// Per-pulse update for the poses already in d_poses: device transform + refit, without host synchronisation.
// Only targets that have moved since the scene was committed are touched (e->moving); the SAH cost of the
// refitted tree comes back asynchronously and is examined at the next call: when it has drifted more than
// RTS_REBUILD_RATIO above the cost the topology had when it was built (targets far from where they were
// clustered), the tree is rebuilt at the new poses instead of refitted.
int bvh_refit(rts_engine *e)
{
    int rc;
    bvh_join(e);
    collect_sah(e, false);
    if (e->n_tris >= 2 && e->sah_at_build > 0 && e->bvh_info.sah_cost > RTS_REBUILD_RATIO * e->sah_at_build) return bvh_build(e);
    uint32_t n_moving = 0;
    for (uint32_t k = 0; k < e->n_targets; k++) n_moving += e->moving[k] ? 1u : 0u;
    if (n_moving == 0) return RTS_OK;
    if (n_moving == e->n_targets || e->n_tris < 4096) {
        if ((rc = bvh_update_world(e))) return rc;
        if ((rc = fit_and_pack(e))) return rc;
    } else {
        if (!e->partial_ready && (rc = prepare_partial(e))) return rc;
        if ((rc = partial_update(e))) return rc;
    }
    if (e->n_tris >= 2) {
        if (!e->bvh_join_pending) {   // (a refit that went to side_bvh has enqueued its read-back there)
        collect_sah(e, true);   // one read-back slot: the previous one has long arrived
        RTS_CUDA(cudaMemcpyAsync(&e->h_rb->sah, e->d_sah, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        cudaEventRecord(e->sah_ev, e->stream);
        e->sah_pending = true;
        }
        // The first refits after a build are examined at once: committing base meshes and then placing the
        // targets (the usual first call) can move them arbitrarily far from where the topology clustered them.
        if (e->refits_since_build < 2) {
            e->refits_since_build++;
            collect_sah(e, true);
            if (e->sah_at_build > 0 && e->bvh_info.sah_cost > RTS_REBUILD_RATIO * e->sah_at_build) {
                e->builder = e->builder_forced;   // the geometry the builder was chosen on was not representative: choose again
                return bvh_build(e);
            }
        }
    }
    return RTS_OK;
}

// Wait for the last refit's SAH cost (diagnostics: rts_scene_bvh_info).
void bvh_sync_info(rts_engine *e) { collect_sah(e, true); }

int bvh_check(rts_engine *e, uint64_t *violations)
{
    const int n = (int)e->n_tris;
    unsigned long long v = 0;
    bvh_join(e);
    if (n >= 2) {
        RTS_CUDA(cudaMemsetAsync(e->d_violations, 0, sizeof(unsigned long long), e->stream));
        { k_check<<<blocks_for(n, 256), 256, 0, e->stream>>>(e->d_parent, e->d_order, e->d_tri_box, e->d_node_box, n,
                                                          e->d_violations); e->launches++; }
        RTS_CUDA(cudaGetLastError());
        RTS_CUDA(cudaMemcpyAsync(&v, e->d_violations, sizeof(v), cudaMemcpyDeviceToHost, e->stream));
        RTS_CUDA(cudaStreamSynchronize(e->stream));
    }
    if (violations) *violations = v;
    return RTS_OK;
}
