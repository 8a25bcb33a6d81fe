// oracle_scene.cpp — scene flattening and the oracle's own simple BVH (median split, fp64 boxes).
// TEST INFRASTRUCTURE (see rts_oracle.h).  The BVH exists only so the oracle finishes on
// 100k–1M-triangle scenes; the definitional truth is the exhaustive search (use_bvh = 0) and
// tests cross-check the two on sub-samples.
#include "oracle_common.h"
#include <algorithm>
#include <numeric>

namespace orc {

void build_scene(Scene &s, const rts_target_mesh *targets, uint32_t n_targets)
{
    s.meshes.clear();
    s.total_tris = 0;
    for (uint32_t i = 0; i < n_targets; i++) {
        Mesh m;
        m.verts = targets[i].verts; m.tris = targets[i].tris; m.normals = targets[i].normals;
        m.n_verts = targets[i].n_verts; m.n_tris = targets[i].n_tris; m.n_normals = targets[i].n_normals;
        m.refl_coeff = targets[i].refl_coeff; m.refr_index = targets[i].refr_index;
        m.tri_offset = s.total_tris;
        s.total_tris += m.n_tris;
        s.meshes.push_back(m);
    }
    s.tri_mesh.resize(s.total_tris);
    for (uint32_t i = 0; i < n_targets; i++)
        for (uint32_t t = 0; t < s.meshes[i].n_tris; t++) s.tri_mesh[s.meshes[i].tri_offset + t] = i;
    s.has_bvh = false;
}

namespace {
struct Box { double lo[3], hi[3]; };

struct Builder {
    Scene &s;
    std::vector<Box> boxes;
    std::vector<double> cent; // 3 per tri
    explicit Builder(Scene &sc) : s(sc) {}

    int32_t build(uint32_t start, uint32_t count)
    {
        BvhNode nd;
        for (int a = 0; a < 3; a++) { nd.lo[a] = INFINITY; nd.hi[a] = -INFINITY; }
        double clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t i = start; i < start + count; i++) {
            const Box &b = boxes[s.order[i]];
            for (int a = 0; a < 3; a++) {
                nd.lo[a] = std::min(nd.lo[a], b.lo[a]);
                nd.hi[a] = std::max(nd.hi[a], b.hi[a]);
                clo[a] = std::min(clo[a], cent[3 * (size_t)s.order[i] + a]);
                chi[a] = std::max(chi[a], cent[3 * (size_t)s.order[i] + a]);
            }
        }
        nd.left = nd.right = -1; nd.start = (int32_t)start; nd.count = 0;
        const int32_t me = (int32_t)s.nodes.size();
        s.nodes.push_back(nd);
        if (count <= 4) {
            s.nodes[me].count = (int32_t)count;
            return me;
        }
        int axis = 0;
        double ext = chi[0] - clo[0];
        for (int a = 1; a < 3; a++)
            if (chi[a] - clo[a] > ext) { ext = chi[a] - clo[a]; axis = a; }
        uint32_t mid = start + count / 2;
        std::nth_element(s.order.begin() + start, s.order.begin() + mid, s.order.begin() + start + count,
                         [&](uint32_t x, uint32_t y) { return cent[3 * (size_t)x + axis] < cent[3 * (size_t)y + axis]; });
        int32_t l = build(start, mid - start);
        int32_t r = build(mid, start + count - mid);
        s.nodes[me].left = l;
        s.nodes[me].right = r;
        return me;
    }
};
} // namespace

void build_bvh(Scene &s)
{
    Builder b(s);
    b.boxes.resize(s.total_tris);
    b.cent.resize(3 * (size_t)s.total_tris);
    for (const Mesh &m : s.meshes) {
        for (uint32_t t = 0; t < m.n_tris; t++) {
            Box bx;
            for (int a = 0; a < 3; a++) { bx.lo[a] = INFINITY; bx.hi[a] = -INFINITY; }
            for (int k = 0; k < 3; k++) {
                const double *v = m.verts + 3 * (size_t)m.tris[3 * (size_t)t + k];
                for (int a = 0; a < 3; a++) { bx.lo[a] = std::min(bx.lo[a], v[a]); bx.hi[a] = std::max(bx.hi[a], v[a]); }
            }
            const uint32_t g = m.tri_offset + t;
            b.boxes[g] = bx;
            for (int a = 0; a < 3; a++) b.cent[3 * (size_t)g + a] = 0.5 * (bx.lo[a] + bx.hi[a]);
        }
    }
    s.order.resize(s.total_tris);
    std::iota(s.order.begin(), s.order.end(), 0u);
    s.nodes.clear();
    s.nodes.reserve(s.total_tris);
    b.build(0, s.total_tris);
    s.has_bvh = true;
}

} // namespace orc
