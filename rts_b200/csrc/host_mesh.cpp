// host_mesh.cpp — pure-host helpers of the C-ABI: receiver-sphere set-up, the reference's three
// mesh generators and its rigid rotation.  Compiled with -ffp-contract=off.
//
// Reference behaviour reproduced (file:line in /root/reference/ray_tracer.cpp):
//   receiver sphere / window   :894-918  (float trig on double angles, as written there)
//   vertex_rotation            :156-170  (FLOAT yaw/pitch/roll; cos/sin evaluated in float)
//   rect_mesh                  :226-297  (8 vertices, 12 triangles, 12 FACE normals)
//   sphere_mesh                :300-426  (subdivided icosahedron; vertices de-duplicated through an
//                                         ordered set => lexicographically sorted; triangles likewise)
//   file_mesh                  :429-504  ("x y z, x y z, x y z," per line; 3 fresh vertices per triangle)
// Flat arrays instead of nested vectors; arithmetic order identical so results match bit for bit.
#include "../../include/rts_b200.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <map>
#include <vector>

extern "C" void rts_rx_sphere_from_desc(const rts_rx_desc *d, rts_rx_sphere *out)
{
    const double az = d->azimuth, el = d->elevation, rho = d->radius;
    // centre = position + rho * (cos el cos az, cos el sin az, sin el), trig in float (:903-905)
    out->centre[0] = d->position[0] + (rho * cosf(el) * cosf(az));
    out->centre[1] = d->position[1] + (rho * cosf(el) * sinf(az));
    out->centre[2] = d->position[2] + (rho * sinf(el));
    // receiver position seen from the centre (:908-910)
    const double ddx = d->position[0] - out->centre[0], ddy = d->position[1] - out->centre[1],
                 ddz = d->position[2] - out->centre[2];
    const double theta0 = atan2f(ddy, ddx);
    const double phi0 = atan2f(ddz, sqrt(ddx * ddx + ddy * ddy));
    out->radius = rho;
    out->min_theta = theta0 - d->theta_span / 2;
    out->max_theta = theta0 + d->theta_span / 2;
    out->min_phi = phi0 - d->phi_span / 2;
    out->max_phi = phi0 + d->phi_span / 2;
}

namespace {

void mul3(const double A[9], const double B[9], double C[9])
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += A[3 * i + k] * B[3 * k + j];
            C[3 * i + j] = acc;
        }
}

void rotate_in_place(double *xyz, size_t n, const double R[9])
{
    for (size_t v = 0; v < n; v++) {
        double *p = xyz + 3 * v;
        double o[3];
        for (int i = 0; i < 3; i++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += R[3 * i + k] * p[k];
            o[i] = acc;
        }
        p[0] = o[0]; p[1] = o[1]; p[2] = o[2];
    }
}

int emit(const std::vector<double> &v, const std::vector<uint32_t> &t, const std::vector<double> &n, double *ov,
         uint32_t *nv, uint32_t *ot, uint32_t *nt, double *on, uint32_t *nn)
{
    if (nv) *nv = (uint32_t)(v.size() / 3);
    if (nt) *nt = (uint32_t)(t.size() / 3);
    if (nn) *nn = (uint32_t)(n.size() / 3);
    if (ov) std::copy(v.begin(), v.end(), ov);
    if (ot) std::copy(t.begin(), t.end(), ot);
    if (on) std::copy(n.begin(), n.end(), on);
    return RTS_OK;
}

} // namespace

extern "C" void rts_rotation_matrix(float yaw, float pitch, float roll, double R[9])
{
    // float-precision cos/sin widened to double, then Rz * (Ry * Rx) in double (:158-162)
    const double cr = std::cos(roll), sr = std::sin(roll), cp = std::cos(pitch), sp = std::sin(pitch),
                 cy = std::cos(yaw), sy = std::sin(yaw);
    const double Rx[9] = {1, 0, 0, 0, cr, -sr, 0, sr, cr};
    const double Ry[9] = {cp, 0, sp, 0, 1, 0, -sp, 0, cp};
    const double Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
    double RyRx[9];
    mul3(Ry, Rx, RyRx);
    mul3(Rz, RyRx, R);
}

extern "C" int rts_rect_mesh(float w, float h, float d, float yaw, float pitch, float roll, double *ov, uint32_t *nv,
                             uint32_t *ot, uint32_t *nt, double *on, uint32_t *nn)
{
    std::vector<double> v(24);
    for (int i = 0; i < 8; i++) {
        const float fx = (i & 4) ? -0.5f : +0.5f;   // vertices 0-3: +w/2, 4-7: -w/2
        const float fy = (i & 1) ? +0.5f : -0.5f;
        const float fz = (i & 2) ? +0.5f : -0.5f;
        v[3 * i] = w * fx; v[3 * i + 1] = h * fy; v[3 * i + 2] = d * fz;   // float products, widened
    }
    static const uint32_t T[36] = {0, 1, 2, 1, 3, 2, 2, 3, 7, 2, 7, 6, 1, 7, 3, 1, 5, 7,
                                   6, 7, 4, 7, 5, 4, 0, 4, 1, 1, 4, 5, 2, 6, 4, 0, 2, 4};
    std::vector<uint32_t> t(T, T + 36);
    double R[9];
    rts_rotation_matrix(yaw, pitch, roll, R);
    rotate_in_place(v.data(), 8, R);
    std::vector<double> fn(36);
    for (int i = 0; i < 12; i++) {
        const double *a = &v[3 * t[3 * i]], *b = &v[3 * t[3 * i + 1]], *c = &v[3 * t[3 * i + 2]];
        const double u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, q[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        double x = (u[1] * q[2] - u[2] * q[1]), y = (u[2] * q[0] - u[0] * q[2]), z = (u[0] * q[1] - u[1] * q[0]);
        const double norm = sqrt(x * x + y * y + z * z);
        fn[3 * i] = x / norm; fn[3 * i + 1] = y / norm; fn[3 * i + 2] = z / norm;
    }
    return emit(v, t, fn, ov, nv, ot, nt, on, nn);
}

extern "C" int rts_sphere_mesh(uint32_t subdivs, float radius, float yaw, float pitch, float roll, double *ov,
                               uint32_t *nv, uint32_t *ot, uint32_t *nt, double *on, uint32_t *nn)
{
    typedef std::array<double, 3> P3;
    const double g = (1 + sqrt(5)) / 2;
    std::vector<P3> v = {{-1, g, 0}, {1, g, 0}, {-1, -g, 0}, {1, -g, 0}, {0, -1, g}, {0, 1, g},
                         {0, -1, -g}, {0, 1, -g}, {g, 0, -1}, {g, 0, 1}, {-g, 0, -1}, {-g, 0, 1}};
    for (P3 &p : v) {
        const double norm = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        p = {p[0] / norm, p[1] / norm, p[2] / norm};
    }
    typedef std::array<uint32_t, 3> I3;
    std::vector<I3> f = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                         {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                         {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    auto mid = [&](uint32_t a, uint32_t b) -> uint32_t { // appended every time, duplicates included (:85-101)
        P3 m = {(v[a][0] + v[b][0]) / 2, (v[a][1] + v[b][1]) / 2, (v[a][2] + v[b][2]) / 2};
        const double norm = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
        m = {m[0] / norm, m[1] / norm, m[2] / norm};
        v.push_back(m);
        return (uint32_t)v.size() - 1;
    };
    for (uint32_t gen = 0; gen < subdivs; gen++) {
        std::vector<I3> f2;
        f2.reserve(f.size() * 4);
        for (const I3 &tri : f) {
            const uint32_t a = mid(tri[0], tri[1]), b = mid(tri[1], tri[2]), c = mid(tri[2], tri[0]);
            f2.push_back({tri[0], a, c});
            f2.push_back({tri[1], b, a});
            f2.push_back({tri[2], c, b});
            f2.push_back({a, b, c});
        }
        f.swap(f2);
    }
    // unique vertices in lexicographic order; std::map gives the rank of each original vertex
    std::map<P3, uint32_t> rank;
    for (const P3 &p : v) rank.emplace(p, 0u);
    std::vector<double> verts;
    verts.reserve(rank.size() * 3);
    uint32_t r = 0;
    for (auto &kv : rank) {
        kv.second = r++;
        verts.insert(verts.end(), kv.first.begin(), kv.first.end());
    }
    double R[9];
    rts_rotation_matrix(yaw, pitch, roll, R);
    rotate_in_place(verts.data(), verts.size() / 3, R);
    std::vector<double> normals = verts; // unit vectors (:407)
    for (I3 &tri : f) tri = {rank[v[tri[0]]], rank[v[tri[1]]], rank[v[tri[2]]]};
    std::sort(f.begin(), f.end());
    f.erase(std::unique(f.begin(), f.end()), f.end());
    std::vector<uint32_t> tris;
    tris.reserve(f.size() * 3);
    for (const I3 &tri : f) tris.insert(tris.end(), tri.begin(), tri.end());
    for (double &x : verts) x *= radius;
    return emit(verts, tris, normals, ov, nv, ot, nt, on, nn);
}

extern "C" int rts_file_mesh(const char *v_file, const char *n_file, float yaw, float pitch, float roll, double *ov,
                             uint32_t *nv, uint32_t *ot, uint32_t *nt, double *on, uint32_t *nn)
{
    if (!v_file || !n_file) return RTS_ERR_ARG;
    auto load = [](const char *path, std::vector<double> &out, long expect) -> long {
        FILE *fp = fopen(path, "r");
        if (!fp) return -1;
        long lines = 0;
        for (int ch; (ch = fgetc(fp)) != EOF;) lines += ch == '\n';
        if (expect >= 0) lines = expect;
        rewind(fp);
        out.assign((size_t)lines * 9, 0.0);
        for (long i = 0; i < lines; i++) {
            double *p = &out[(size_t)i * 9];
            if (fscanf(fp, "%lf %lf %lf, %lf %lf %lf, %lf %lf %lf,\n", p, p + 1, p + 2, p + 3, p + 4, p + 5, p + 6, p + 7,
                       p + 8) == EOF) {
                fclose(fp);
                return -2;
            }
        }
        fclose(fp);
        return lines;
    };
    std::vector<double> verts, normals;
    const long ntri = load(v_file, verts, -1);
    if (ntri < 0) return RTS_ERR_ARG;
    if (load(n_file, normals, ntri) < 0) return RTS_ERR_ARG;
    std::vector<uint32_t> tris((size_t)ntri * 3);
    for (size_t i = 0; i < tris.size(); i++) tris[i] = (uint32_t)i;
    double R[9];
    rts_rotation_matrix(yaw, pitch, roll, R);
    rotate_in_place(verts.data(), verts.size() / 3, R);
    rotate_in_place(normals.data(), normals.size() / 3, R);
    return emit(verts, tris, normals, ov, nv, ot, nt, on, nn);
}
