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
__global__ void k_transform(const double *__restrict__ base, double *__restrict__ world,
                            const uint32_t *__restrict__ owner, const rts_pose *__restrict__ poses, uint32_t n,
                            int translate)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
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
                            float *__restrict__ tri_box, unsigned *__restrict__ scene_box, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    if (i < n) {
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

__global__ void k_leaf_of_tri(const uint32_t *__restrict__ order, uint32_t *__restrict__ leaf_of_tri, uint32_t n)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) leaf_of_tri[order[p]] = p;
}

// Leaf-ordered triangle records: one thread per leaf position, 5 x 128-bit stores.
__global__ void k_tri_records(const double *__restrict__ world, const uint32_t *__restrict__ tris,
                              const uint32_t *__restrict__ tri_target, const uint32_t *__restrict__ t_vert_off,
                              const uint32_t *__restrict__ order, TriRec *__restrict__ rec, uint32_t n)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t id = order[p];
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

// Bottom-up fit: the second thread to arrive at a node unions its children and carries on.
__global__ void k_fit(const int2 *__restrict__ children, const int32_t *__restrict__ parent,
                      const uint32_t *__restrict__ order, const float *tri_box, float *node_box,
                      uint32_t *__restrict__ flags, int n)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int cur = parent[(n - 1) + p];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(flags + cur, 1u) == 0u) return;
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

__global__ void k_pack(const int2 *__restrict__ children, const int2 *__restrict__ range,
                       const uint32_t *__restrict__ order, const float *__restrict__ tri_box,
                       const float *__restrict__ node_box, BvhNode *__restrict__ nodes, int n, double *sah, int leaf_max)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
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
    if ((threadIdx.x & 31) == 0 && area > 0.0) atomicAdd(sah, area);
    if (i >= n - 1) return;
    const int2 ch = children[i];
    BvhNode nd;
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
        }
    }
    nd.c0x = cc[0][0]; nd.c0y = cc[0][1]; nd.h0x = hh[0][0]; nd.h0y = hh[0][1];
    nd.c1x = cc[1][0]; nd.c1y = cc[1][1]; nd.h1x = hh[1][0]; nd.h1y = hh[1][1];
    nd.c0z = cc[0][2]; nd.c1z = cc[1][2]; nd.h0z = hh[0][2]; nd.h1z = hh[1][2];
    nd.ref0 = refs[0]; nd.ref1 = refs[1]; nd.pad[0] = nd.pad[1] = 0;
    float4 *dst = reinterpret_cast<float4 *>(nodes + i);
    const float4 *s4 = reinterpret_cast<const float4 *>(&nd);
    dst[0] = s4[0]; dst[1] = s4[1]; dst[2] = s4[2]; dst[3] = s4[3];
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
                     (void **)&e->d_nodes, (void **)&e->d_trirec, (void **)&e->d_cub_temp, (void **)&e->d_violations, (void **)&e->d_sah};
    for (void **p : ptrs) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
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
    if ((rc = dalloc(&e->d_trirec, T))) return rc;
    if ((rc = dalloc(&e->d_violations, (size_t)1))) return rc;
    if ((rc = dalloc(&e->d_sah, (size_t)1))) return rc;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, e->d_morton, e->d_morton_sorted, e->d_order_in, e->d_order, (int)T, 0,
                                    63, e->stream);
    e->cub_temp_bytes = bytes;
    if (e->d_cub_temp) { cudaFree(e->d_cub_temp); e->d_cub_temp = nullptr; }
    RTS_CUDA(cudaMalloc(&e->d_cub_temp, bytes ? bytes : 16));
    return RTS_OK;
}

int bvh_update_world(rts_engine *e)
{
    const unsigned bs = 256;
    if (e->n_verts)
        { k_transform<<<blocks_for(e->n_verts, bs), bs, 0, e->stream>>>(e->d_base_verts, e->d_world_verts, e->d_vert_target,
                                                                     e->d_poses, e->n_verts, 1); e->launches++; }
    if (e->n_normals)
        { k_transform<<<blocks_for(e->n_normals, bs), bs, 0, e->stream>>>(e->d_base_normals, e->d_world_normals,
                                                                       e->d_norm_target, e->d_poses, e->n_normals, 0); e->launches++; }
    // scene box reset: lo = +inf, hi = -inf in ordered encoding
    static const unsigned init_box[6] = {0xff800000u, 0xff800000u, 0xff800000u, 0x007fffffu, 0x007fffffu, 0x007fffffu};
    RTS_CUDA(cudaMemcpyAsync(e->d_scene_box, init_box, sizeof(init_box), cudaMemcpyHostToDevice, e->stream));
    if (e->n_tris)
        { k_tri_boxes<<<blocks_for(e->n_tris, bs), bs, 0, e->stream>>>(e->d_world_verts, e->d_tris, e->d_tri_target,
                                                                    e->d_t_vert_off, e->d_tri_box,
                                                                    (unsigned *)e->d_scene_box, e->n_tris); e->launches++; }
    RTS_CUDA(cudaGetLastError());
    return RTS_OK;
}

static int fit_and_pack(rts_engine *e)
{
    const unsigned bs = 256;
    const int n = (int)e->n_tris;
    if (n == 0) { e->root_ref = 0; return RTS_OK; }
    { k_tri_records<<<blocks_for(n, bs), bs, 0, e->stream>>>(e->d_world_verts, e->d_tris, e->d_tri_target, e->d_t_vert_off,
                                                          e->d_order, e->d_trirec, n); e->launches++; }
    if (n >= 2) {
        RTS_CUDA(cudaMemsetAsync(e->d_fit_flags, 0, sizeof(uint32_t) * (size_t)n, e->stream));
        { k_fit<<<blocks_for(n, bs), bs, 0, e->stream>>>(e->d_children, e->d_parent, e->d_order, e->d_tri_box, e->d_node_box,
                                                      e->d_fit_flags, n); e->launches++; }
        RTS_CUDA(cudaMemsetAsync(e->d_sah, 0, sizeof(double), e->stream));
        { k_pack<<<blocks_for(n - 1, bs), bs, 0, e->stream>>>(e->d_children, e->d_range, e->d_order, e->d_tri_box,
                                                           e->d_node_box, e->d_nodes, n, e->d_sah, e->leaf_max); e->launches++; }
    }
    RTS_CUDA(cudaGetLastError());
    e->root_ref = n <= e->leaf_max ? ~((0 << 3) | (n - 1)) : 0;
    return RTS_OK;
}

static int read_scene_box(rts_engine *e)
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
        const float m = fmaxf(fabsf(bi.scene_lo[a]), fabsf(bi.scene_hi[a]));
        e->scene_abs[a] = (e->n_tris && m < 3.0e38f) ? m : 0.f;
    }
    return RTS_OK;
}

int bvh_build(rts_engine *e)
{
    const unsigned bs = 256;
    const int n = (int)e->n_tris;
    cudaEventRecord(e->ev[4], e->stream);
    int rc = bvh_update_world(e);
    if (rc) return rc;
    if (n > 0) {
        { k_morton<<<blocks_for(n, bs), bs, 0, e->stream>>>(e->d_tri_box, (const unsigned *)e->d_scene_box, e->d_morton,
                                                         e->d_order_in, n); e->launches++; }
        size_t bytes = e->cub_temp_bytes;
        RTS_CUDA(cub::DeviceRadixSort::SortPairs(e->d_cub_temp, bytes, e->d_morton, e->d_morton_sorted, e->d_order_in,
                                                 e->d_order, n, 0, 63, e->stream));
        { k_leaf_of_tri<<<blocks_for(n, bs), bs, 0, e->stream>>>(e->d_order, e->d_leaf_of_tri, n); e->launches++; }
        if (n >= 2)
            { k_hierarchy<<<blocks_for(n - 1, bs), bs, 0, e->stream>>>(e->d_morton_sorted, n, e->d_children, e->d_range,
                                                                    e->d_parent); e->launches++; }
        RTS_CUDA(cudaGetLastError());
    }
    rc = fit_and_pack(e);
    if (rc) return rc;
    cudaEventRecord(e->ev[5], e->stream);
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev[4], e->ev[5]);
    if ((rc = read_scene_box(e))) return rc;
    rts_bvh_info &bi = e->bvh_info;
    bi.n_tris = e->n_tris;
    bi.n_nodes = n >= 2 ? n - 1 : 0;
    bi.root_is_leaf = e->root_ref < 0;
    bi.max_leaf = e->leaf_max;
    bi.ms_build = ms;
    double sah = 0;
    if (n >= 2) RTS_CUDA(cudaMemcpy(&sah, e->d_sah, sizeof(double), cudaMemcpyDeviceToHost));
    bi.sah_cost = sah;
    e->sah_at_build = sah;
    e->builds++;
    return RTS_OK;
}

// Refit; when the tree's SAH cost has drifted more than RTS_REBUILD_RATIO above the cost it had when
// its topology was built (targets moved far from where they were clustered), rebuild instead.
int bvh_refit(rts_engine *e)
{
    int rc = bvh_update_world(e);
    if (rc) return rc;
    if ((rc = fit_and_pack(e))) return rc;
    if ((rc = read_scene_box(e))) return rc;
    if (e->n_tris >= 2) {
        double sah = 0;
        RTS_CUDA(cudaMemcpyAsync(&sah, e->d_sah, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        RTS_CUDA(cudaStreamSynchronize(e->stream));
        e->bvh_info.sah_cost = sah;
        if (sah > RTS_REBUILD_RATIO * e->sah_at_build) return bvh_build(e);
    }
    return RTS_OK;
}

int bvh_check(rts_engine *e, uint64_t *violations)
{
    const int n = (int)e->n_tris;
    unsigned long long v = 0;
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
