// bvh_width_study.cpp — planning aid (CPU only, not part of the library or the oracle): how many node fetches and box
// tests would first reflections off a C4-like terrain cost in a 2-, 4- or 8-wide BVH?  DESIGN.md §9 item 1.
//
// A 1000 x 500-cell height field (4 m cells, ~30 m relief), the benchmark's transmitter position, one reflected ray per
// sampled facet (direction = mirror of the incoming direction about the facet normal, origin on the facet).  A median
// split binary tree with leaves of <= 2 triangles is collapsed to width W by absorbing log2(W) levels per node; traversal
// is closest-hit with a stack, children tested with exact slabs and visited near to far.  Counted per ray: wide nodes
// fetched (the dependent steps of the GPU loop), child boxes tested, triangles tested.
//   g++ -O2 -std=c++17 -fopenmp tools/bvh_width_study.cpp -o /tmp/bvh_width_study && /tmp/bvh_width_study
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <vector>

struct V { double x, y, z; };
static V sub(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static double dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V norm(V a) { double l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }

struct Box { double lo[3], hi[3]; };
struct Node { Box box; int left = -1, right = -1, start = 0, count = 0; };   // leaf when count > 0

static std::vector<V> P;               // vertices
static std::vector<int> T;             // 3 per triangle
static std::vector<int> order;
static std::vector<Box> tbox;
static std::vector<V> cent;
static std::vector<Node> nodes;

static double height(double x, double y)
{
    return 12.0 * std::sin(x * 0.011) * std::cos(y * 0.017) + 8.0 * std::sin(x * 0.031 + 1.3) * std::sin(y * 0.023) + 5.0 * std::sin(x * 0.083 + y * 0.057) +
           3.0 * std::cos(y * 0.19 - x * 0.13) + 2.0 * std::sin(x * 0.41) * std::cos(y * 0.37);
}

static int build(int start, int count)
{
    Node nd;
    for (int a = 0; a < 3; a++) { nd.box.lo[a] = 1e300; nd.box.hi[a] = -1e300; }
    double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
    for (int i = start; i < start + count; i++) {
        const Box &b = tbox[order[i]];
        const double c[3] = {cent[order[i]].x, cent[order[i]].y, cent[order[i]].z};
        for (int a = 0; a < 3; a++) {
            nd.box.lo[a] = std::min(nd.box.lo[a], b.lo[a]); nd.box.hi[a] = std::max(nd.box.hi[a], b.hi[a]);
            clo[a] = std::min(clo[a], c[a]); chi[a] = std::max(chi[a], c[a]);
        }
    }
    nd.start = start;
    const int me = (int)nodes.size();
    nodes.push_back(nd);
    if (count <= 2) { nodes[me].count = count; return me; }
    int axis = 0;
    for (int a = 1; a < 3; a++) if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
    const int mid = start + count / 2;
    std::nth_element(order.begin() + start, order.begin() + mid, order.begin() + start + count, [&](int p, int q) {
        const double cp[3] = {cent[p].x, cent[p].y, cent[p].z}, cq[3] = {cent[q].x, cent[q].y, cent[q].z};
        return cp[axis] < cq[axis];
    });
    const int l = build(start, mid - start), r = build(mid, start + count - mid);
    nodes[me].left = l; nodes[me].right = r;
    return me;
}

static bool slab(const Box &b, V o, V inv, double tmax, double &tn)
{
    const double oo[3] = {o.x, o.y, o.z}, ii[3] = {inv.x, inv.y, inv.z};
    double t0 = 0, t1 = tmax;
    for (int a = 0; a < 3; a++) {
        double x0 = (b.lo[a] - oo[a]) * ii[a], x1 = (b.hi[a] - oo[a]) * ii[a];
        if (x0 > x1) std::swap(x0, x1);
        t0 = std::max(t0, x0); t1 = std::min(t1, x1);
    }
    tn = t0;
    return t0 <= t1;
}

struct Counts { double fetches = 0, boxes = 0, tris = 0, hits = 0; };

// children of wide node `n`: descend `levels` binary levels, stopping at leaves
static void gather(int n, int levels, std::vector<int> &out)
{
    if (levels == 0 || nodes[n].count > 0) { out.push_back(n); return; }
    gather(nodes[n].left, levels - 1, out);
    gather(nodes[n].right, levels - 1, out);
}

static void trace(V o, V d, int levels, Counts &c)
{
    const V inv = {1 / d.x, 1 / d.y, 1 / d.z};
    double best = 1e30;
    std::vector<int> stack = {0}, kids;
    std::vector<std::pair<double, int>> hit;
    while (!stack.empty()) {
        const int n = stack.back(); stack.pop_back();
        if (nodes[n].count > 0) {
            for (int k = 0; k < nodes[n].count; k++) {
                c.tris++;
                const int t = order[nodes[n].start + k];
                const V p0 = P[T[3 * t]], p1 = P[T[3 * t + 1]], p2 = P[T[3 * t + 2]];
                const V e0 = sub(p1, p0), e1 = sub(p0, p2), nn = cross(e1, e0);
                const double den = dot(nn, d);
                const V e2 = {(p0.x - o.x) / den, (p0.y - o.y) / den, (p0.z - o.z) / den};
                const double tt = dot(nn, e2);
                if (!(tt > 0.005 && tt < best)) continue;
                const V i = cross(d, e2);
                const double beta = dot(i, e1), gamma = dot(i, e0);
                if (beta >= 0 && gamma >= 0 && beta + gamma <= 1) best = tt;
            }
            continue;
        }
        c.fetches++;                       // one wide node: all its child boxes arrive with one dependent fetch
        kids.clear(); hit.clear();
        gather(nodes[n].left, levels - 1, kids);
        gather(nodes[n].right, levels - 1, kids);
        for (int k : kids) {
            c.boxes++;
            double tn;
            if (slab(nodes[k].box, o, inv, best, tn)) hit.push_back({tn, k});
        }
        std::sort(hit.begin(), hit.end(), [](auto &a, auto &b) { return a.first > b.first; });   // nearest on top
        for (auto &h : hit) stack.push_back(h.second);
    }
    if (best < 1e29) c.hits++;
}

int main()
{
    const int cx = 1000, cy = 500;
    const double cell = 4.0;
    for (int j = 0; j <= cy; j++)
        for (int i = 0; i <= cx; i++) { const double x = i * cell, y = (j - cy / 2) * cell; P.push_back({x, y, height(x, y)}); }
    auto vid = [&](int i, int j) { return j * (cx + 1) + i; };
    for (int j = 0; j < cy; j++)
        for (int i = 0; i < cx; i++) {
            const int a = vid(i, j), b = vid(i + 1, j), c = vid(i + 1, j + 1), d = vid(i, j + 1);
            T.insert(T.end(), {a, b, c, a, c, d});
        }
    const int nt = (int)T.size() / 3;
    tbox.resize(nt); cent.resize(nt); order.resize(nt);
    std::iota(order.begin(), order.end(), 0);
    for (int t = 0; t < nt; t++) {
        Box b;
        for (int a = 0; a < 3; a++) { b.lo[a] = 1e300; b.hi[a] = -1e300; }
        for (int k = 0; k < 3; k++) {
            const V p = P[T[3 * t + k]];
            const double v[3] = {p.x, p.y, p.z};
            for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], v[a]); b.hi[a] = std::max(b.hi[a], v[a]); }
        }
        tbox[t] = b;
        cent[t] = {0.5 * (b.lo[0] + b.hi[0]), 0.5 * (b.lo[1] + b.hi[1]), 0.5 * (b.lo[2] + b.hi[2])};
    }
    nodes.reserve(2 * nt);
    build(0, nt);
    std::printf("%d triangles, %zu binary nodes\n", nt, nodes.size());
    const V tx = {-3000.0, 0.0, 800.0};
    // one reflected ray per sampled facet that faces the transmitter
    std::vector<std::pair<V, V>> rays;
    for (int t = 0; t < nt; t += 37) {
        const V p0 = P[T[3 * t]], p1 = P[T[3 * t + 1]], p2 = P[T[3 * t + 2]];
        const V c = {(p0.x + p1.x + p2.x) / 3, (p0.y + p1.y + p2.y) / 3, (p0.z + p1.z + p2.z) / 3};
        V n = norm(cross(sub(p1, p0), sub(p2, p0)));
        const V in = norm(sub(c, tx));
        if (dot(n, in) > 0) n = {-n.x, -n.y, -n.z};
        const double k = 2 * dot(n, in);
        if (-dot(n, in) < 0.02) continue;                 // grazing or facing away: not illuminated
        rays.push_back({c, norm(V{in.x - k * n.x, in.y - k * n.y, in.z - k * n.z})});
    }
    std::printf("%zu reflected rays\n", rays.size());
    std::printf("%-6s %14s %14s %14s %10s\n", "width", "fetches/ray", "box tests/ray", "tri tests/ray", "hit %");
    for (int levels = 1; levels <= 4; levels++) {
        Counts tot;
#pragma omp parallel
        {
            Counts c;
#pragma omp for schedule(dynamic, 256)
            for (long i = 0; i < (long)rays.size(); i++) trace(rays[i].first, rays[i].second, levels, c);
#pragma omp critical
            { tot.fetches += c.fetches; tot.boxes += c.boxes; tot.tris += c.tris; tot.hits += c.hits; }
        }
        const double n = (double)rays.size();
        std::printf("%-6d %14.2f %14.2f %14.2f %10.2f\n", 1 << levels, tot.fetches / n, tot.boxes / n, tot.tris / n, 100.0 * tot.hits / n);
    }
    return 0;
}
