// ploc.cuh — agglomerative BVH topology by parallel locally-ordered clustering (included inside bvh.cu's namespace).
//
// The Morton LBVH (k_hierarchy) splits by code bits only; PLOC (Meister & Bittner, "Parallel Locally-Ordered
// Clustering for Bounding Volume Hierarchy Construction", TVCG 2018) starts from the same Morton order but builds
// bottom-up: every cluster looks RADIUS neighbours left and right for the partner with the smallest merged surface
// area, mutual nearest neighbours merge, the array is compacted, repeat until one cluster is left.  The result is
// close to a full-sweep SAH tree at a few milliseconds for a million triangles, and it feeds the very same arrays
// the refit / packing / traversal code already uses: children[], parent[], range[], order[] with the root at 0.
//
// Outputs are made to look exactly like k_hierarchy's: internal node ids run 0 .. n-2 with the root at 0 (ids are
// handed out in reverse creation order), leaf references are ~position where positions are the left-to-right leaf
// order of the final tree (so every subtree owns a contiguous position range, which the leaf collapsing of k_pack
// and the leaf-ordered triangle records rely on).

#define RTS_PLOC_RADIUS 16

struct PlocBox { float lo[3], hi[3]; };

__device__ __forceinline__ float ploc_area(const PlocBox &a, const PlocBox &b)
{
    const float dx = fmaxf(a.hi[0], b.hi[0]) - fminf(a.lo[0], b.lo[0]);
    const float dy = fmaxf(a.hi[1], b.hi[1]) - fminf(a.lo[1], b.lo[1]);
    const float dz = fmaxf(a.hi[2], b.hi[2]) - fminf(a.lo[2], b.lo[2]);
    return dx * dy + dy * dz + dz * dx;
}

// initial clusters: one per triangle in Morton order
__global__ void k_ploc_init(const uint32_t *__restrict__ order0, const float *__restrict__ tri_box, int n, int *__restrict__ ref,
                            PlocBox *__restrict__ box)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ref[i] = ~i;
    const float *b = tri_box + 6 * (size_t)order0[i];
    PlocBox B;
#pragma unroll
    for (int a = 0; a < 3; a++) { B.lo[a] = b[a]; B.hi[a] = b[3 + a]; }
    box[i] = B;
}

// nearest neighbour within the window (ties to the lower index, so the relation is deterministic)
__global__ void k_ploc_nn(const PlocBox *__restrict__ box, int m, int *__restrict__ nn)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const PlocBox me = box[i];
    float best = CUDART_INF_F;
    int arg = -1;
    const int lo = max(0, i - RTS_PLOC_RADIUS), hi = min(m - 1, i + RTS_PLOC_RADIUS);
    for (int j = lo; j <= hi; j++) {
        if (j == i) continue;
        const float a = ploc_area(me, box[j]);
        if (a < best) { best = a; arg = j; }
    }
    nn[i] = arg;
}

// flags: keep[i] = cluster i survives at its place (alone or as the merged pair), merged[i] = it is the left partner
__global__ void k_ploc_flags(const int *__restrict__ nn, int m, int *__restrict__ keep, int *__restrict__ merged)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    keep[i] = (mutual && j < i) ? 0 : 1;
    merged[i] = (mutual && i < j) ? 1 : 0;
}

// apply the merges and compact.  Node ids are handed out downwards from `next_id` (exclusive scan of merged gives the
// rank within this round), so the last merge of the build — the root — receives id 0.
__global__ void k_ploc_merge(const int *__restrict__ nn, const int *__restrict__ keep_scan, const int *__restrict__ merged_scan,
                             const int *__restrict__ keep, const int *__restrict__ merged, int m, int next_id,
                             const int *__restrict__ ref_in, const PlocBox *__restrict__ box_in, int *__restrict__ ref_out,
                             PlocBox *__restrict__ box_out, int2 *__restrict__ children, int32_t *__restrict__ parent_node,
                             int32_t *__restrict__ parent_leaf0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m || !keep[i]) return;
    const int at = keep_scan[i];
    if (!merged[i]) {
        ref_out[at] = ref_in[i];
        box_out[at] = box_in[i];
        return;
    }
    const int j = nn[i];
    const int id = next_id - merged_scan[i];
    const int ra = ref_in[i], rb = ref_in[j];
    children[id] = make_int2(ra, rb);
    if (ra >= 0) parent_node[ra] = id; else parent_leaf0[~ra] = id;
    if (rb >= 0) parent_node[rb] = id; else parent_leaf0[~rb] = id;
    const PlocBox A = box_in[i], B = box_in[j];
    PlocBox U;
#pragma unroll
    for (int a = 0; a < 3; a++) { U.lo[a] = fminf(A.lo[a], B.lo[a]); U.hi[a] = fmaxf(A.hi[a], B.hi[a]); }
    ref_out[at] = id;
    box_out[at] = U;
}

// leaf counts bottom-up (second arrival continues), in the numbering of the initial Morton positions
__global__ void k_ploc_counts(const int2 *__restrict__ children, const int32_t *__restrict__ parent_node,
                              const int32_t *__restrict__ parent_leaf0, int n, uint32_t *__restrict__ flags, int *__restrict__ count)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int cur = parent_leaf0[p];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(flags + cur, 1u) == 0u) return;
        const int2 ch = children[cur];
        const int a = ch.x >= 0 ? __ldcg(count + ch.x) : 1, b = ch.y >= 0 ? __ldcg(count + ch.y) : 1;
        __stcg(count + cur, a + b);
        cur = parent_node[cur];
    }
}

// left-to-right position of every leaf in the final tree: the leaves left of it are the left siblings' subtrees of
// the ancestors it hangs under as a right child
__global__ void k_ploc_positions(const int2 *__restrict__ children, const int32_t *__restrict__ parent_node,
                                 const int32_t *__restrict__ parent_leaf0, const int *__restrict__ count, int n,
                                 const uint32_t *__restrict__ order0, uint32_t *__restrict__ order, int *__restrict__ pos_of0)
{
    const int p0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (p0 >= n) return;
    int pos = 0, from = ~p0, cur = parent_leaf0[p0];
    while (cur >= 0) {
        const int2 ch = children[cur];
        if (ch.y == from) pos += ch.x >= 0 ? count[ch.x] : 1;
        from = cur;
        cur = parent_node[cur];
    }
    pos_of0[p0] = pos;
    order[pos] = order0[p0];
}

// position range of every internal node: its leftmost leaf's position and its leaf count (leaf references are
// still initial Morton positions here)
__global__ void k_ploc_ranges(const int2 *__restrict__ children, const int *__restrict__ pos_of0, const int *__restrict__ count,
                              int n, int2 *__restrict__ range)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int ref = i;
    while (ref >= 0) ref = children[ref].x;
    const int lo = pos_of0[~ref];
    range[i] = make_int2(lo, lo + count[i] - 1);
}

// relabel leaf references to final positions and fill parent[] in the layout k_hierarchy produces
__global__ void k_ploc_finish(int2 *__restrict__ children, const int32_t *__restrict__ parent_node,
                              const int32_t *__restrict__ parent_leaf0, const int *__restrict__ pos_of0, int n,
                              int32_t *__restrict__ parent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) {
        int2 ch = children[i];
        if (ch.x < 0) ch.x = ~pos_of0[~ch.x];
        if (ch.y < 0) ch.y = ~pos_of0[~ch.y];
        children[i] = ch;
        parent[i] = parent_node[i];
    }
    if (i < n) parent[(n - 1) + pos_of0[i]] = parent_leaf0[i];
}
