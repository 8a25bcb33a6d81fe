// oracle_common.h — shared declarations of the CPU oracle (test infrastructure; see rts_oracle.h).
#pragma once
#include "rts_oracle.h"
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

struct D3 { double x, y, z; };
struct F3 { float x, y, z; };

// --- double3 helpers, evaluation order as in the reference helpers
//     (ray_tracer.cu:72-129, triangle_mesh.cu:39-118, normal_shader.cu:48-115) ---
static inline D3 d3(double x, double y, double z) { return D3{x, y, z}; }
static inline D3 add(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline D3 sub(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline D3 scale(double a, D3 b) { return d3(a * b.x, a * b.y, a * b.z); }
static inline D3 cross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double magsq(D3 a) { return (a.x * a.x + a.y * a.y + a.z * a.z); }
static inline double length(D3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
static inline D3 normalised(D3 a) { double n = length(a); return d3(a.x / n, a.y / n, a.z / n); }
// normalise_float3: fp64 normalise, then narrow (ray_tracer.cu:125-129, normal_shader.cu:96-100)
static inline F3 normalise_float3(double x, double y, double z)
{
    double n = length(d3(x, y, z));
    return F3{(float)(x / n), (float)(y / n), (float)(z / n)};
}

struct Mesh {            // one target, world coordinates
    const double *verts; const uint32_t *tris; const double *normals;
    uint32_t n_verts, n_tris, n_normals;
    double refl_coeff, refr_index;
    uint32_t tri_offset;  // global id of local triangle 0
};

struct BvhNode {          // oracle's own acceleration structure (not the product's)
    double lo[3], hi[3];
    int32_t left, right;  // children; leaf when count > 0
    int32_t start, count;
};

struct Scene {
    std::vector<Mesh> meshes;
    uint32_t total_tris = 0;
    std::vector<uint32_t> tri_mesh;      // global tri -> mesh index
    // BVH (optional)
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> order;         // leaf order -> global tri id
    bool has_bvh = false;
};

void build_scene(Scene &s, const rts_target_mesh *targets, uint32_t n_targets);
void build_bvh(Scene &s);

} // namespace orc
