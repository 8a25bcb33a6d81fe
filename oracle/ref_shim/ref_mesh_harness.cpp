// ref_mesh_harness.cpp — TEST INFRASTRUCTURE.  The reference's own host-side mesh helpers (ray_tracer.cpp: getMidPoint,
// triangle_area, matrix_multiply, matrix_transpose, vertex_rotation, rect_mesh, sphere_mesh, file_mesh — STL-only free
// functions) compiled UNMODIFIED: oracle/Makefile cuts the text between "void getMidPoint(" and the "/* Main RTS
// function */" banner out of /root/reference/ray_tracer.cpp into oracle/_ref/ref_mesh_extract.inc at build time (the
// repository holds no reference source; oracle/_ref/ is git-ignored) and this file includes it.  The extern "C" wrappers
// below use the two-call protocol of rts_oracle.h's orc_*_mesh so that tests can bit-compare the three implementations:
// reference code, oracle restatement (oracle_mesh.cpp), library (rts_b200/csrc/host_mesh.cpp).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <iomanip>
#include <iterator>
#include <sstream>
#include <set>
#include <string>
#include <initializer_list>
#include <limits>
#include <stdexcept>
#include <vector>

using std::string;
using std::vector;

#include "../_ref/ref_mesh_extract.inc"

namespace {
int emit(const vector<vector<double>> &v, const vector<vector<unsigned int>> &t, const vector<vector<double>> &n, double *verts,
         uint32_t *n_verts, uint32_t *tris, uint32_t *n_tris, double *normals, uint32_t *n_normals)
{
    *n_verts = (uint32_t)v.size(); *n_tris = (uint32_t)t.size(); *n_normals = (uint32_t)n.size();
    if (verts) for (size_t i = 0; i < v.size(); i++) for (int k = 0; k < 3; k++) verts[3 * i + k] = v[i][k];
    if (tris) for (size_t i = 0; i < t.size(); i++) for (int k = 0; k < 3; k++) tris[3 * i + k] = t[i][k];
    if (normals) for (size_t i = 0; i < n.size(); i++) for (int k = 0; k < 3; k++) normals[3 * i + k] = n[i][k];
    return 0;
}
}

extern "C" {
int refm_rect_mesh(float w, float h, float d, float yaw, float pitch, float roll, double *verts, uint32_t *n_verts, uint32_t *tris,
                   uint32_t *n_tris, double *normals, uint32_t *n_normals)
{
    vector<vector<double>> v, n;
    vector<vector<unsigned int>> t;
    rect_mesh(w, h, d, v, t, n, yaw, pitch, roll);
    return emit(v, t, n, verts, n_verts, tris, n_tris, normals, n_normals);
}
int refm_sphere_mesh(uint32_t subdivs, float radius, float yaw, float pitch, float roll, double *verts, uint32_t *n_verts, uint32_t *tris,
                     uint32_t *n_tris, double *normals, uint32_t *n_normals)
{
    vector<vector<double>> v, n;
    vector<vector<unsigned int>> t;
    unsigned int num = 0;
    sphere_mesh(subdivs, radius, v, t, n, yaw, pitch, roll, num);
    return emit(v, t, n, verts, n_verts, tris, n_tris, normals, n_normals);
}
int refm_file_mesh(const char *v_file, const char *n_file, float yaw, float pitch, float roll, double *verts, uint32_t *n_verts,
                   uint32_t *tris, uint32_t *n_tris, double *normals, uint32_t *n_normals)
{
    vector<vector<double>> v, n;
    vector<vector<unsigned int>> t;
    file_mesh(v_file, n_file, v, t, n, yaw, pitch, roll);
    return emit(v, t, n, verts, n_verts, tris, n_tris, normals, n_normals);
}
void refm_vertex_rotation(double *xyz, uint32_t n, float yaw, float pitch, float roll)
{
    vector<vector<double>> v(n, vector<double>(3));
    for (uint32_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) v[i][k] = xyz[3 * i + k];
    v = vertex_rotation(v, yaw, pitch, roll);
    for (uint32_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) xyz[3 * i + k] = v[i][k];
}
}
