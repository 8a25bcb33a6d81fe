// oracle_mesh.cpp — the reference's mesh generators and rigid rotation, restated.
// TEST INFRASTRUCTURE (see rts_oracle.h).
//   vertex_rotation /root/reference/ray_tracer.cpp:156-170  (float angles: std::cos(float) is cosf)
//   rect_mesh       /root/reference/ray_tracer.cpp:226-297
//   sphere_mesh     /root/reference/ray_tracer.cpp:300-426  (getMidPoint :85-101)
//   file_mesh       /root/reference/ray_tracer.cpp:429-504
#include "oracle_common.h"
#include <algorithm>
#include <cstdio>
#include <set>
#include <vector>

namespace {
typedef std::vector<std::vector<double>> Mat;

Mat mat_mul(const Mat &A, const Mat &B) // ray_tracer.cpp:120-137
{
    Mat C(A.size(), std::vector<double>(B[0].size(), 0));
    for (size_t i = 0; i < A.size(); i++)
        for (size_t j = 0; j < B[0].size(); j++) {
            C[i][j] = 0;
            for (size_t k = 0; k < B.size(); k++) C[i][j] += A[i][k] * B[k][j];
        }
    return C;
}

Mat rotation_total(float yaw, float pitch, float roll) // ray_tracer.cpp:158-162
{
    Mat Rx = {{1, 0, 0}, {0, std::cos(roll), -std::sin(roll)}, {0, std::sin(roll), std::cos(roll)}};
    Mat Ry = {{std::cos(pitch), 0, std::sin(pitch)}, {0, 1, 0}, {-std::sin(pitch), 0, std::cos(pitch)}};
    Mat Rz = {{std::cos(yaw), -std::sin(yaw), 0}, {std::sin(yaw), std::cos(yaw), 0}, {0, 0, 1}};
    return mat_mul(Rz, mat_mul(Ry, Rx));
}

void rotate_rows(Mat &v, float yaw, float pitch, float roll) // (R * v^T)^T, ray_tracer.cpp:166
{
    Mat R = rotation_total(yaw, pitch, roll);
    for (auto &row : v) {
        double out[3];
        for (int i = 0; i < 3; i++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += R[i][k] * row[k];
            out[i] = acc;
        }
        row[0] = out[0]; row[1] = out[1]; row[2] = out[2];
    }
}

int emit(const Mat &verts, const std::vector<std::vector<unsigned>> &tris, const Mat &normals, double *ov,
         uint32_t *nv, uint32_t *ot, uint32_t *nt, double *on, uint32_t *nn)
{
    if (nv) *nv = (uint32_t)verts.size();
    if (nt) *nt = (uint32_t)tris.size();
    if (nn) *nn = (uint32_t)normals.size();
    if (ov) for (size_t i = 0; i < verts.size(); i++) for (int k = 0; k < 3; k++) ov[3 * i + k] = verts[i][k];
    if (ot) for (size_t i = 0; i < tris.size(); i++) for (int k = 0; k < 3; k++) ot[3 * i + k] = tris[i][k];
    if (on) for (size_t i = 0; i < normals.size(); i++) for (int k = 0; k < 3; k++) on[3 * i + k] = normals[i][k];
    return 0;
}
} // namespace

extern "C" void orc_vertex_rotation(double *xyz, uint32_t n, float yaw, float pitch, float roll)
{
    Mat v(n, std::vector<double>(3));
    for (uint32_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) v[i][k] = xyz[3 * (size_t)i + k];
    rotate_rows(v, yaw, pitch, roll);
    for (uint32_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) xyz[3 * (size_t)i + k] = v[i][k];
}

extern "C" int orc_rect_mesh(float w, float h, float d, float yaw, float pitch, float roll, double *ov, uint32_t *nv,
                             uint32_t *ot, uint32_t *nt, double *on, uint32_t *nn)
{
    Mat vertices(8, std::vector<double>(3));
    const float sx[8] = {+0.5f, +0.5f, +0.5f, +0.5f, -0.5f, -0.5f, -0.5f, -0.5f};
    const float sy[8] = {-0.5f, +0.5f, -0.5f, +0.5f, -0.5f, +0.5f, -0.5f, +0.5f};
    const float sz[8] = {-0.5f, -0.5f, +0.5f, +0.5f, -0.5f, -0.5f, +0.5f, +0.5f};
    for (int i = 0; i < 8; i++) {
        vertices[i][0] = w * sx[i]; // float product, then widened (ray_tracer.cpp:235-242)
        vertices[i][1] = h * sy[i];
        vertices[i][2] = d * sz[i];
    }
    const unsigned T[12][3] = {{0, 1, 2}, {1, 3, 2}, {2, 3, 7}, {2, 7, 6}, {1, 7, 3}, {1, 5, 7},
                               {6, 7, 4}, {7, 5, 4}, {0, 4, 1}, {1, 4, 5}, {2, 6, 4}, {0, 2, 4}};
    std::vector<std::vector<unsigned>> tris(12, std::vector<unsigned>(3));
    for (int i = 0; i < 12; i++) for (int k = 0; k < 3; k++) tris[i][k] = T[i][k];
    rotate_rows(vertices, yaw, pitch, roll);
    Mat face(12, std::vector<double>(3, 0)); // ray_tracer.cpp:269-292
    for (int i = 0; i < 12; i++) {
        double v1[3], v2[3];
        for (int k = 0; k < 3; k++) {
            v1[k] = vertices[tris[i][1]][k] - vertices[tris[i][0]][k];
            v2[k] = vertices[tris[i][2]][k] - vertices[tris[i][0]][k];
        }
        face[i][0] = (v1[1] * v2[2] - v1[2] * v2[1]);
        face[i][1] = (v1[2] * v2[0] - v1[0] * v2[2]);
        face[i][2] = (v1[0] * v2[1] - v1[1] * v2[0]);
        double norm = sqrt(face[i][0] * face[i][0] + face[i][1] * face[i][1] + face[i][2] * face[i][2]);
        face[i][0] = face[i][0] / norm;
        face[i][1] = face[i][1] / norm;
        face[i][2] = face[i][2] / norm;
    }
    return emit(vertices, tris, face, ov, nv, ot, nt, on, nn);
}

extern "C" int orc_sphere_mesh(uint32_t n, float radius, float yaw, float pitch, float roll, double *ov, uint32_t *nv,
                               uint32_t *ot, uint32_t *nt, double *on, uint32_t *nn)
{
    double t = (1 + sqrt(5)) / 2;
    Mat v = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
             {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    for (auto &p : v) {
        double norm = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        p[0] = p[0] / norm; p[1] = p[1] / norm; p[2] = p[2] / norm;
    }
    std::vector<std::vector<unsigned>> f = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                                            {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                                            {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    auto midpoint = [&](int a, int b) { // ray_tracer.cpp:85-101
        std::vector<double> pm(3, 0);
        pm[0] = (v[a][0] + v[b][0]) / 2;
        pm[1] = (v[a][1] + v[b][1]) / 2;
        pm[2] = (v[a][2] + v[b][2]) / 2;
        double norm = sqrt(pm[0] * pm[0] + pm[1] * pm[1] + pm[2] * pm[2]);
        pm[0] = pm[0] / norm; pm[1] = pm[1] / norm; pm[2] = pm[2] / norm;
        v.push_back(pm);
    };
    for (unsigned gen = 0; gen < n; gen++) {
        std::vector<std::vector<unsigned>> f_(f.size() * 4, std::vector<unsigned>(3, 0));
        for (size_t i = 0; i < f.size(); i++) {
            int tri[3] = {(int)f[i][0], (int)f[i][1], (int)f[i][2]};
            int a = (int)v.size(); midpoint(tri[0], tri[1]);
            int b = (int)v.size(); midpoint(tri[1], tri[2]);
            int c = (int)v.size(); midpoint(tri[2], tri[0]);
            int nfc[4][3] = {{tri[0], a, c}, {tri[1], b, a}, {tri[2], c, b}, {a, b, c}};
            for (int j = 0; j < 4; j++) for (int k = 0; k < 3; k++) f_[4 * i + j][k] = (unsigned)nfc[j][k];
        }
        f = f_;
    }
    // duplicate removal through an ordered set (ray_tracer.cpp:394-401): vertices end up sorted
    std::set<std::vector<double>> v_unique(v.begin(), v.end());
    Mat verts(v_unique.begin(), v_unique.end());
    std::vector<int> ix(v.size());
    for (size_t i = 0; i < v.size(); i++)
        ix[i] = (int)(std::lower_bound(verts.begin(), verts.end(), v[i]) - verts.begin());
    rotate_rows(verts, yaw, pitch, roll);
    Mat normals = verts; // unit vectors (ray_tracer.cpp:407)
    for (auto &tri : f) for (int k = 0; k < 3; k++) tri[k] = (unsigned)ix[tri[k]];
    std::set<std::vector<unsigned>> f_unique(f.begin(), f.end());
    std::vector<std::vector<unsigned>> tris(f_unique.begin(), f_unique.end());
    for (auto &p : verts) { p[0] *= radius; p[1] *= radius; p[2] *= radius; }
    return emit(verts, tris, normals, ov, nv, ot, nt, on, nn);
}

extern "C" int orc_file_mesh(const char *v_file, const char *n_file, float yaw, float pitch, float roll, double *ov,
                             uint32_t *nv, uint32_t *ot, uint32_t *nt, double *on, uint32_t *nn)
{
    FILE *fp = fopen(v_file, "r");
    if (!fp) return -1;
    unsigned lines = 0;
    for (int ch; (ch = fgetc(fp)) != EOF;) lines += ch == '\n';
    rewind(fp);
    Mat vertices(lines * 3, std::vector<double>(3)), normals(lines * 3, std::vector<double>(3));
    std::vector<std::vector<unsigned>> tris(lines, std::vector<unsigned>(3));
    for (unsigned i = 0; i < lines; i++) for (unsigned k = 0; k < 3; k++) tris[i][k] = i * 3 + k;
    auto read9 = [&](FILE *f, Mat &m) {
        for (unsigned i = 0; i < lines; i++)
            if (fscanf(f, "%lf %lf %lf, %lf %lf %lf, %lf %lf %lf,\n", &m[3 * i][0], &m[3 * i][1], &m[3 * i][2],
                       &m[3 * i + 1][0], &m[3 * i + 1][1], &m[3 * i + 1][2], &m[3 * i + 2][0], &m[3 * i + 2][1],
                       &m[3 * i + 2][2]) == EOF)
                return -1;
        return 0;
    };
    int rc = read9(fp, vertices);
    fclose(fp);
    if (rc) return -2;
    rotate_rows(vertices, yaw, pitch, roll);
    fp = fopen(n_file, "r");
    if (!fp) return -1;
    rc = read9(fp, normals);
    fclose(fp);
    if (rc) return -2;
    rotate_rows(normals, yaw, pitch, roll);
    return emit(vertices, tris, normals, ov, nv, ot, nt, on, nn);
}
