// oracle.cpp — TEST INFRASTRUCTURE ONLY (oracle tier B): deterministic CPU restatement of the reference's
// hot path.  Imported / linked only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
//
// Every function cites the reference lines it follows.  Arithmetic: IEEE float, -ffp-contract=off, glm's
// scalar operation order.  The only deliberate difference from the reference is the random
// number source: the reference's five racy std::default_random_engine objects (main.cpp:57,
// pathTracing.cpp:33,106,113,149) are replaced by Philox4x32-10 keyed on (seed; pixel, sample, depth, slot),
// the same convention the GPU uses, so that "same sample seeds" exists (SURVEY §0 finding 5).
//
// Pinned by tests/test_oracle_vs_ref.py against oracle/_ref/libref.so (the unmodified reference objects).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <omp.h>
#include <vector>

namespace
{
// ---------------------------------------------------------------------------------------------- vec3 (glm order)
struct V3
{
    float x = 0, y = 0, z = 0;
    V3() {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(V3 v) { return std::sqrt(dot(v, v)); }
inline V3 normalize(V3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline V3 reflect(V3 I, V3 N) { return I - N * dot(N, I) * 2.0f; }
inline V3 refract(V3 I, V3 N, float eta)
{
    const float d = dot(N, I);
    const float k = 1.0f - eta * eta * (1.0f - d * d);
    return (k >= 0.0f) ? (eta * I - (eta * d + std::sqrt(k)) * N) : V3();
}

const float INF = 114514.0f;  // bvh.h:5
const float PI = 3.1415926f;  // pathtracing.h:11
const float P_RR = 0.8f;      // pathtracing.h:12
enum { DIFFUSE = 0, SPECULAR = 1, TRANSMISSION = 2, INVALID = 3 }; // ray.h:5-8

struct Tri // triangle.h:9-26 (material by index instead of by name)
{
    V3 v[3], vn[3];
    float vt[3][2];
    V3 normal, center;
    double area = 0;
    int mtl = 0;
    bool emissive = false;
    int src = 0; // input (OBJ) index
};

struct Mat // material.h:11-33
{
    V3 Kd, Ks, Tr;
    float Ns = 1, Ni = 1;
    int texture = -1;
    bool emissive = false;
    V3 radiance;
    double area = 0;
    std::vector<Tri> tris; // light triangles, OBJ order, cumulative area (scene.cpp:199-205)
};

struct Tex
{
    int rows = 0, cols = 0;
    std::vector<uint8_t> bgr;
};

struct Node // bvh.h:16-22
{
    int left = -1, right = -1, index = 0, num = 0;
    V3 AA, BB;
};

struct Hit // bvh.h:7-15
{
    bool is_hit = false;
    float distance = INF;
    V3 hitpoint, direction, pn;
    int tri = -1;
    bool emissive = false; // triangle.is_emissive of the record (false for a default record)
};
} // namespace

struct orc_scene
{
    std::vector<Tri> tris;
    std::vector<Mat> mats;
    std::vector<int> lights; // material index per <light>, XML order
    std::vector<Tex> tex;
    std::vector<Node> nodes; // pre-order, nodes[0] = root
    std::vector<uint32_t> key; // A.4 tie key per post-build triangle (used by the pruned counter walk only)
    V3 eye, llc, horizontal, vertical;
    int W = 0, H = 0;
};

namespace
{
// ---------------------------------------------------------------------------------------------- triangle.cpp
double calAera(const Tri &t) // triangle.cpp:3-10
{
    double a = length(t.v[1] - t.v[0]), b = length(t.v[2] - t.v[0]), c = length(t.v[2] - t.v[1]);
    double cos_c = (a * a + b * b - c * c) / (2 * a * b);
    double sin_c = std::sqrt(1 - std::pow(cos_c, 2));
    return a * b * sin_c / 2;
}

// triangle.cpp:12-29: least squares [v0 v1 v2; 1 1 1] b = [p; 1] in double by column-pivoted Householder QR
// (the algorithm behind Eigen's colPivHouseholderQr().solve(); Eigen itself is un-vendored / unpinned).
V3 findBaryCor(const Tri &t, V3 p)
{
    double A[4][3] = {{t.v[0].x, t.v[1].x, t.v[2].x}, {t.v[0].y, t.v[1].y, t.v[2].y}, {t.v[0].z, t.v[1].z, t.v[2].z}, {1, 1, 1}};
    double b[4] = {p.x, p.y, p.z, 1};
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < 3; k++)
    {
        int best = k;
        double bn = -1;
        for (int j = k; j < 3; j++)
        {
            double s = 0;
            for (int i = k; i < 4; i++)
                s += A[i][j] * A[i][j];
            if (s > bn)
                bn = s, best = j;
        }
        if (best != k)
        {
            for (int i = 0; i < 4; i++)
                std::swap(A[i][k], A[i][best]);
            std::swap(perm[k], perm[best]);
        }
        double norm = std::sqrt(bn);
        if (norm == 0)
            continue;
        double alpha = (A[k][k] > 0) ? -norm : norm;
        double w[4] = {0, 0, 0, 0};
        for (int i = k; i < 4; i++)
            w[i] = A[i][k];
        w[k] -= alpha;
        double wtw = 0;
        for (int i = k; i < 4; i++)
            wtw += w[i] * w[i];
        if (wtw == 0)
            continue;
        double beta = 2 / wtw;
        for (int j = k; j < 3; j++)
        {
            double s = 0;
            for (int i = k; i < 4; i++)
                s += w[i] * A[i][j];
            s *= beta;
            for (int i = k; i < 4; i++)
                A[i][j] -= s * w[i];
        }
        double s = 0;
        for (int i = k; i < 4; i++)
            s += w[i] * b[i];
        s *= beta;
        for (int i = k; i < 4; i++)
            b[i] -= s * w[i];
    }
    double y[3], r[3];
    for (int k = 2; k >= 0; k--)
    {
        double s = b[k];
        for (int j = k + 1; j < 3; j++)
            s -= A[k][j] * y[j];
        y[k] = (A[k][k] != 0) ? s / A[k][k] : 0;
    }
    for (int k = 0; k < 3; k++)
        r[perm[k]] = y[k];
    return V3((float)r[0], (float)r[1], (float)r[2]);
}

// ---------------------------------------------------------------------------------------------- bvh.cpp build
float triMin(const Tri &t, int a)
{
    const float c[3] = {(&t.v[0].x)[a], (&t.v[1].x)[a], (&t.v[2].x)[a]};
    return gmin(c[0], gmin(c[1], c[2]));
}
float triMax(const Tri &t, int a)
{
    const float c[3] = {(&t.v[0].x)[a], (&t.v[1].x)[a], (&t.v[2].x)[a]};
    return gmax(c[0], gmax(c[1], c[2]));
}

void sortRange(std::vector<Tri> &tris, int l, int r, int axis) // bvh.cpp:3-14,56-60
{
    if (axis == 0)
        std::sort(tris.begin() + l, tris.begin() + r + 1, [](const Tri &a, const Tri &b) { return a.center.x < b.center.x; });
    if (axis == 1)
        std::sort(tris.begin() + l, tris.begin() + r + 1, [](const Tri &a, const Tri &b) { return a.center.y < b.center.y; });
    if (axis == 2)
        std::sort(tris.begin() + l, tris.begin() + r + 1, [](const Tri &a, const Tri &b) { return a.center.z < b.center.z; });
}

int buildBVH(orc_scene &s, int l, int r, int leaf_num) // bvh.cpp:16-144
{
    if (l > r)
        return -1;
    std::vector<Tri> &tris = s.tris;
    const int me = (int)s.nodes.size();
    s.nodes.push_back(Node());
    Node nd;
    nd.AA = V3(1145141919.f, 1145141919.f, 1145141919.f);
    nd.BB = V3(-1145141919.f, -1145141919.f, -1145141919.f);
    for (int i = l; i <= r; i++) // :25-41
    {
        nd.AA.x = gmin(nd.AA.x, triMin(tris[i], 0) - 0.001f);
        nd.AA.y = gmin(nd.AA.y, triMin(tris[i], 1) - 0.001f);
        nd.AA.z = gmin(nd.AA.z, triMin(tris[i], 2) - 0.001f);
        nd.BB.x = gmax(nd.BB.x, triMax(tris[i], 0) + 0.001f);
        nd.BB.y = gmax(nd.BB.y, triMax(tris[i], 1) + 0.001f);
        nd.BB.z = gmax(nd.BB.z, triMax(tris[i], 2) + 0.001f);
    }
    if ((r - l + 1) <= leaf_num) // :43-48
    {
        nd.num = r - l + 1;
        nd.index = l;
        s.nodes[me] = nd;
        return me;
    }
    float Cost = INF; // :49-51
    int Axis = 0;
    int Split = (l + r) / 2;
    const int n = r - l + 1;
    for (int axis = 0; axis < 3; axis++)
    {
        sortRange(tris, l, r, axis);
        std::vector<V3> lmax(n, V3(-INF, -INF, -INF)), lmin(n, V3(INF, INF, INF)); // :62-63
        for (int i = l; i <= r; i++) // :65-77
        {
            const int bias = (i == l) ? 0 : 1;
            const V3 pmax = lmax[i - l - bias], pmin = lmin[i - l - bias];
            lmax[i - l] = V3(gmax(pmax.x, triMax(tris[i], 0)), gmax(pmax.y, triMax(tris[i], 1)), gmax(pmax.z, triMax(tris[i], 2)));
            lmin[i - l] = V3(gmin(pmin.x, triMin(tris[i], 0)), gmin(pmin.y, triMin(tris[i], 1)), gmin(pmin.z, triMin(tris[i], 2)));
        }
        std::vector<V3> rmax(n, V3(-INF, -INF, -INF)), rmin(n, V3(INF, INF, INF)); // :79-80
        for (int i = r; i >= l; i--) // :82-94
        {
            const int bias = (i == r) ? 0 : 1;
            const V3 pmax = rmax[i - l + bias], pmin = rmin[i - l + bias];
            rmax[i - l] = V3(gmax(pmax.x, triMax(tris[i], 0)), gmax(pmax.y, triMax(tris[i], 1)), gmax(pmax.z, triMax(tris[i], 2)));
            rmin[i - l] = V3(gmin(pmin.x, triMin(tris[i], 0)), gmin(pmin.y, triMin(tris[i], 1)), gmin(pmin.z, triMin(tris[i], 2)));
        }
        float cost = INF; // :96-123
        int split = l;
        for (int i = l; i <= r - 1; i++)
        {
            float xl = lmax[i - l].x - lmin[i - l].x, yl = lmax[i - l].y - lmin[i - l].y, zl = lmax[i - l].z - lmin[i - l].z;
            float la = 2.0 * ((xl * yl) + (xl * zl) + (yl * zl));
            float lc = la * (i - l + 1);
            xl = rmax[i + 1 - l].x - rmin[i + 1 - l].x, yl = rmax[i + 1 - l].y - rmin[i + 1 - l].y, zl = rmax[i + 1 - l].z - rmin[i + 1 - l].z;
            float ra = 2.0 * ((xl * yl) + (xl * zl) + (yl * zl));
            float rc = ra * (r - i);
            float total = lc + rc;
            if (total < cost)
                cost = total, split = i;
        }
        if (cost < Cost) // :125-130
            Cost = cost, Axis = axis, Split = split;
    }
    sortRange(tris, l, r, Axis); // :133-138
    nd.left = buildBVH(s, l, Split, leaf_num);
    nd.right = buildBVH(s, Split + 1, r, leaf_num);
    s.nodes[me] = nd;
    return me;
}

// ---------------------------------------------------------------------------------------------- bvh.cpp traverse
struct Counters
{
    uint64_t box = 0, tri = 0;
};

float interactAABB(V3 S, V3 d, V3 AA, V3 BB) // bvh.cpp:231-245
{
    V3 inv((float)(1.0 / d.x), (float)(1.0 / d.y), (float)(1.0 / d.z));
    V3 in = (BB - S) * inv;
    V3 out = (AA - S) * inv;
    V3 tmax(gmax(in.x, out.x), gmax(in.y, out.y), gmax(in.z, out.z));
    V3 tmin(gmin(in.x, out.x), gmin(in.y, out.y), gmin(in.z, out.z));
    float t1 = gmin(tmax.x, gmin(tmax.y, tmax.z));
    float t0 = gmax(tmin.x, gmax(tmin.y, tmin.z));
    return (t1 >= t0) ? ((t0 > 0.0) ? (t0) : (t1)) : (-1);
}

bool interactTriangle(const Tri &tr, V3 S, V3 d, float &t_out, V3 &P_out) // bvh.cpp:177-209
{
    V3 p1 = tr.v[0], p2 = tr.v[1], p3 = tr.v[2], N = tr.normal;
    if (std::fabs(dot(N, d)) < 0.00001f)
        return false;
    float t = (dot(p1 - S, N)) / dot(d, N);
    if (t < 0.0005f)
        return false;
    V3 P = S + d * t;
    V3 c1 = cross(p2 - p1, P - p1), c2 = cross(p3 - p2, P - p2), c3 = cross(p1 - p3, P - p3);
    double dir1 = dot(c1, N), dir2 = dot(c2, N), dir3 = dot(c3, N);
    bool r1 = dir1 > 0 && dir2 > 0 && dir3 > 0;
    bool r2 = dir1 < 0 && dir2 < 0 && dir3 < 0;
    if (r1 || r2)
    {
        t_out = t;
        P_out = P;
        return true;
    }
    return false;
}

Hit interactBVHNode(const orc_scene &s, V3 S, V3 d, int l, int r, Counters *cnt) // bvh.cpp:211-229
{
    Hit res;
    for (int i = l; i <= r; i++)
    {
        float t;
        V3 P;
        if (cnt)
            cnt->tri++;
        if (interactTriangle(s.tris[i], S, d, t, P))
        {
            if ((t == res.distance && s.tris[i].emissive) || (t < res.distance))
            {
                res.is_hit = true;
                res.hitpoint = P;
                res.distance = t;
                res.direction = d;
                res.tri = i;
                res.emissive = s.tris[i].emissive;
                V3 bc = findBaryCor(s.tris[i], P);
                const Tri &T = s.tris[i];
                res.pn = normalize((T.vn[0] * bc.x) + (T.vn[1] * bc.y) + (T.vn[2] * bc.z));
            }
        }
    }
    return res;
}

Hit traverseBVH(const orc_scene &s, V3 S, V3 d, int node, Counters *cnt) // bvh.cpp:146-175
{
    if (node < 0)
        return Hit();
    const Node &nd = s.nodes[node];
    if (nd.num > 0)
        return interactBVHNode(s, S, d, nd.index, nd.index + nd.num - 1, cnt);
    float d1 = INF, d2 = INF;
    if (nd.left >= 0)
    {
        d1 = interactAABB(S, d, s.nodes[nd.left].AA, s.nodes[nd.left].BB);
        if (cnt)
            cnt->box++;
    }
    if (nd.right >= 0)
    {
        d2 = interactAABB(S, d, s.nodes[nd.right].AA, s.nodes[nd.right].BB);
        if (cnt)
            cnt->box++;
    }
    Hit r1, r2;
    if (d1 > 0)
        r1 = traverseBVH(s, S, d, nd.left, cnt);
    if (d2 > 0)
        r2 = traverseBVH(s, S, d, nd.right, cnt);
    if (r1.distance == r2.distance) // :168-172 (the else binds to the inner if)
    {
        if (r1.emissive)
            return r1;
        else
            return r2;
    }
    return r1.distance < r2.distance ? r1 : r2;
}

Hit traceRoot(const orc_scene &s, V3 S, V3 d, Counters *cnt = nullptr)
{
    return traverseBVH(s, S, d, s.nodes.empty() ? -1 : 0, cnt);
}

// Ordered, distance-pruned binary walk of the same topology with the A.4 tie key: the accounting basis
// for the roofline's algorithmic bytes.  Not reference code; cross-checked against traceRoot.
int tracePruned(const orc_scene &s, V3 S, V3 d, Counters &cnt)
{
    if (s.nodes.empty())
        return -1;
    float best = INF;
    uint32_t bestKey = 0x7FFFFFFFu;
    int bestId = -1;
    struct E
    {
        int node;
        float t0;
    };
    std::vector<E> stack;
    stack.push_back({0, -1.f});
    const V3 inv((float)(1.0 / d.x), (float)(1.0 / d.y), (float)(1.0 / d.z));
    auto box = [&](const Node &c, float &t0) {
        cnt.box++;
        V3 in = (c.BB - S) * inv, out = (c.AA - S) * inv;
        float t1 = gmin(gmax(in.x, out.x), gmin(gmax(in.y, out.y), gmax(in.z, out.z)));
        t0 = gmax(gmin(in.x, out.x), gmax(gmin(in.y, out.y), gmin(in.z, out.z)));
        return (t1 >= t0) && (((t0 > 0.0f) ? t0 : t1) > 0.0f);
    };
    while (!stack.empty())
    {
        E e = stack.back();
        stack.pop_back();
        if (e.t0 > best)
            continue;
        const Node &nd = s.nodes[e.node];
        if (nd.num > 0)
        {
            for (int i = nd.index; i < nd.index + nd.num; i++)
            {
                float t;
                V3 P;
                cnt.tri++;
                if (interactTriangle(s.tris[i], S, d, t, P) && (t < best || (t == best && s.key[i] > bestKey)))
                    best = t, bestKey = s.key[i], bestId = i;
            }
            continue;
        }
        float tl, tr;
        bool hl = box(s.nodes[nd.left], tl) && !(tl > best);
        bool hr = box(s.nodes[nd.right], tr) && !(tr > best);
        if (hl && hr)
        {
            if (tr < tl)
                stack.push_back({nd.left, tl}), stack.push_back({nd.right, tr});
            else
                stack.push_back({nd.right, tr}), stack.push_back({nd.left, tl});
        }
        else if (hl)
            stack.push_back({nd.left, tl});
        else if (hr)
            stack.push_back({nd.right, tr});
    }
    return bestId;
}

// ---------------------------------------------------------------------------------------------- Philox4x32-10
void philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4])
{
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
    for (int r = 0; r < 10; r++)
    {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

// RNG convention shared with the GPU: counter = (pixel, sample, depth, slot >> 1), key = seed;
// the two doubles of a block come from words (0,1) and (2,3): u = ((hi << 21) | (lo >> 11)) * 2^-53 in [0,1).
enum Slot
{
    S_JITTER_X = 0,
    S_JITTER_Y = 1,
    S_RR = 2,
    S_FRESNEL = 3,
    S_LOBE = 4,
    S_PHI = 5,
    S_THETA = 6,
    S_LIGHT0 = 8 // + 4*light: pick, bary1, bary2, bary3
};

struct Rng
{
    uint64_t seed;
    uint32_t pixel, sample;
    double u(uint32_t depth, uint32_t slot) const
    {
        uint32_t c[4] = {pixel, sample, depth, slot >> 1}, k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
        philox(c, k, o);
        const uint32_t hi = o[(slot & 1) * 2], lo = o[(slot & 1) * 2 + 1];
        return (double)(((uint64_t)hi << 21) | (uint64_t)(lo >> 11)) * (1.0 / 9007199254740992.0);
    }
};

// ---------------------------------------------------------------------------------------------- pathTracing.cpp
V3 Sample(V3 direction, int ray_type, double Ns, double u_phi, double u_theta) // pathTracing.cpp:111-145
{
    double phi = u_phi * 2 * PI;
    double theta;
    if (ray_type == DIFFUSE)
        theta = std::asin(std::sqrt(u_theta));
    else
        theta = std::acos(std::pow(u_theta, (double)1 / (Ns + 1)));
    V3 sample((float)(std::sin(theta) * std::cos(phi)), (float)std::cos(theta), (float)(std::sin(theta) * std::sin(phi)));
    V3 front;
    if (std::fabs(direction.x) > std::fabs(direction.y))
        front = normalize(V3(direction.z, 0, -direction.x));
    else
        front = normalize(V3(0, -direction.z, direction.y));
    V3 right = cross(direction, front);
    return normalize((right * sample.x) + (direction * sample.y) + (front * sample.z));
}

struct NextRay
{
    V3 dir;
    int type = INVALID;
};

NextRay nextRay(const orc_scene &s, const Hit &rec, V3 ray_direction, const Rng &rng, int depth) // pathTracing.cpp:147-209
{
    const Mat &m = s.mats[s.tris[rec.tri].mtl];
    NextRay out;
    if (m.Ni > 1)
    {
        double n1, n2;
        double cos_in = dot(ray_direction, rec.pn);
        V3 normal;
        if (cos_in > 0)
            normal = -rec.pn, n1 = m.Ni, n2 = 1.0;
        else
            normal = rec.pn, n1 = 1.0, n2 = m.Ni;
        double rf0 = std::pow((n1 - n2) / (n1 + n2), 2);
        double fresnel = rf0 + (1.0f - rf0) * std::pow(1.0f - std::abs(cos_in), 5);
        if (fresnel < rng.u(depth, S_FRESNEL))
        {
            V3 refr = refract(ray_direction, normal, (float)(n1 / n2));
            if (!(refr.x == 0 && refr.y == 0 && refr.z == 0))
            {
                out.dir = refr, out.type = TRANSMISSION;
                return out;
            }
            out.dir = reflect(ray_direction, normal), out.type = SPECULAR;
            return out;
        }
    }
    double Kd_len = length(m.Kd), Ks_len = length(m.Ks);
    double kd = Kd_len / (Kd_len + Ks_len), ks = Ks_len / (Kd_len + Ks_len);
    double p = rng.u(depth, S_LOBE);
    if (p < kd)
    {
        out.dir = Sample(rec.pn, DIFFUSE, m.Ns, rng.u(depth, S_PHI), rng.u(depth, S_THETA));
        out.type = DIFFUSE;
    }
    else if (m.Ns > 1 && p < kd + ks)
    {
        V3 reflect_ray = reflect(ray_direction, rec.pn);
        out.dir = Sample(reflect_ray, SPECULAR, m.Ns, rng.u(depth, S_PHI), rng.u(depth, S_THETA));
        out.type = SPECULAR;
    }
    return out;
}

V3 shade(const orc_scene &s, const Hit &rec, V3 wi, const Rng &rng, int depth, int max_depth, uint64_t *counts) // pathTracing.cpp:3-102
{
    V3 L_dir, L_indir;
    const Tri &tri = s.tris[rec.tri];
    if (tri.emissive) // :9-12
        return s.mats[tri.mtl].radiance;
    const Mat &m = s.mats[tri.mtl];
    V3 Kd;
    if (m.texture >= 0) // :17-26
    {
        V3 bc = findBaryCor(tri, rec.hitpoint);
        double col = tri.vt[0][0] * bc.x + tri.vt[1][0] * bc.y + tri.vt[2][0] * bc.z;
        double row = tri.vt[0][1] * bc.x + tri.vt[1][1] * bc.y + tri.vt[2][1] * bc.z;
        double irow = row - std::floor(row), icol = col - std::floor(col);
        const Tex &tx = s.tex[m.texture];
        int r = irow * tx.rows, c = icol * tx.cols;
        // row - floor(row) is exactly 1.0 for a tiny negative row (e.g. -1e-20): the reference then reads one row /
        // column past the image (:24, undefined behaviour); restatement and GPU both clamp to the last texel
        r = std::min(r, tx.rows - 1), c = std::min(c, tx.cols - 1);
        const uint8_t *px = &tx.bgr[((size_t)r * tx.cols + c) * 3];
        Kd.x = (double)px[2] / 255, Kd.y = (double)px[1] / 255, Kd.z = (double)px[0] / 255;
    }
    else
        Kd = m.Kd;

    // direct illumination :33-75.  u1 is a function-local static distribution initialised on first use with
    // the FIRST light's area and reused for every light (:38).
    const double first_area = s.lights.empty() ? 0.0 : s.mats[s.lights[0]].area;
    for (size_t li = 0; li < s.lights.size(); li++)
    {
        const Mat &lm = s.mats[s.lights[li]];
        double total_area = lm.area;
        double rnd = rng.u(depth, S_LIGHT0 + 4 * li) * first_area;
        for (const Tri &lt : lm.tris)
        {
            if (rnd < lt.area)
            {
                double rnd1 = rng.u(depth, S_LIGHT0 + 4 * li + 1), rnd2 = rng.u(depth, S_LIGHT0 + 4 * li + 2),
                       rnd3 = rng.u(depth, S_LIGHT0 + 4 * li + 3);
                float p1 = rnd1 / (rnd1 + rnd2 + rnd3), p2 = rnd2 / (rnd1 + rnd2 + rnd3), p3 = rnd3 / (rnd1 + rnd2 + rnd3);
                V3 light_p = lt.v[0] * p1 + lt.v[1] * p2 + lt.v[2] * p3;
                V3 light_n = normalize(lt.vn[0] * p1 + lt.vn[1] * p2 + lt.vn[2] * p3);
                V3 wo = normalize(light_p - rec.hitpoint);
                Hit rec_sample = traceRoot(s, rec.hitpoint, wo);
                if (counts)
                    counts[1]++;
                // :54-58 visibility = the closest hit's material is the light's material (by name; a miss has "")
                bool visibility = rec_sample.tri >= 0 && s.tris[rec_sample.tri].mtl == lt.mtl;
                if (visibility && dot(wo, rec.pn) > 0)
                {
                    float pdf_light = double(1) / total_area;
                    float cos_theta_p = std::abs(dot(wo, light_n));
                    float cos_theta = std::abs(dot(wo, rec.pn) / length(rec.pn));
                    V3 radiance = lm.radiance;
                    V3 diff = light_p - rec.hitpoint;
                    V3 intensity = radiance * cos_theta_p * cos_theta / dot(diff, diff) / pdf_light;
                    V3 h = normalize((wi + wo) * 0.5f);
                    double cos_alpha = std::fmax((double)dot(rec.pn, h), 0.0);
                    L_dir = L_dir + intensity * (Kd / PI + m.Ks * (m.Ns + 2.0f) * (float)std::pow(cos_alpha, (double)m.Ns) / (2.0f * PI));
                }
                break;
            }
        }
    }

    // indirect :78-99
    const bool may_bounce = (max_depth == 0) || (depth + 1 < max_depth);
    if (may_bounce && rng.u(depth, S_RR) < (double)P_RR) // RR(): :104-109
    {
        NextRay r = nextRay(s, rec, -wi, rng, depth);
        if (r.type != INVALID) // the reference also traces the INVALID ray and discards the result (:81-82)
        {
            Hit ret = traceRoot(s, rec.hitpoint, r.dir);
            if (counts)
                counts[0]++;
            if (ret.is_hit)
            {
                if (r.type == TRANSMISSION)
                    L_indir = L_indir + m.Tr * (shade(s, ret, -r.dir, rng, depth + 1, max_depth, counts) / P_RR);
                else if (!ret.emissive) // DIFFUSE and SPECULAR both weight by Kd (:87-94)
                    L_indir = L_indir + Kd * (shade(s, ret, -r.dir, rng, depth + 1, max_depth, counts) / P_RR);
            }
        }
    }
    return L_dir + L_indir;
}

void primaryRay(const orc_scene &s, int i, int j, const Rng &rng, V3 &S, V3 &d) // main.cpp:88-95, camera.cpp:19-28
{
    double x = double(j) / double(s.W - 1.0);
    double y = double(s.H - i) / double(s.H - 1.0);
    x += (rng.u(0, S_JITTER_X) - 0.5f) / double(s.W);
    y += (rng.u(0, S_JITTER_Y) - 0.5f) / double(s.H);
    const float sx = (float)x, sy = (float)y;
    S = s.eye;
    d = normalize(s.llc + sx * s.horizontal + sy * s.vertical - s.eye);
}
} // namespace

extern "C"
{
orc_scene *orc_scene_create(int32_t n, const float *v9, const float *vn9, const float *vt6, const int32_t *mtl,
                            int32_t n_materials, const orc_material *materials, int32_t n_lights,
                            const int32_t *light_mtl, const float *light_radiance3, int32_t n_textures,
                            const orc_texture *textures, const float *eye, const float *lookat, const float *up,
                            float fovy_f, int32_t width, int32_t height, int32_t leaf_num)
{
    orc_scene *s = new orc_scene();
    s->W = width, s->H = height;
    s->mats.resize(n_materials);
    for (int i = 0; i < n_materials; i++)
    {
        const orc_material &o = materials[i];
        Mat &m = s->mats[i];
        m.Kd = V3(o.Kd[0], o.Kd[1], o.Kd[2]), m.Ks = V3(o.Ks[0], o.Ks[1], o.Ks[2]), m.Tr = V3(o.Tr[0], o.Tr[1], o.Tr[2]);
        m.Ns = o.Ns, m.Ni = o.Ni, m.texture = o.texture;
    }
    for (int l = 0; l < n_lights; l++) // scene.cpp:50-52
    {
        s->lights.push_back(light_mtl[l]);
        s->mats[light_mtl[l]].emissive = true;
        s->mats[light_mtl[l]].radiance = V3(light_radiance3[3 * l], light_radiance3[3 * l + 1], light_radiance3[3 * l + 2]);
    }
    for (int i = 0; i < n_textures; i++)
    {
        Tex t;
        t.rows = textures[i].rows, t.cols = textures[i].cols;
        t.bgr.assign(textures[i].bgr, textures[i].bgr + (size_t)t.rows * t.cols * 3);
        s->tex.push_back(std::move(t));
    }
    // camera.cpp:3-17
    {
        double fovy = fovy_f; // scene.cpp:16: stof -> double
        double aspect = (double)width / (double)height;
        double theta = fovy * 0.01745329251994329576923690768489;
        double h = std::tan(theta / 2);
        float vh = 2.0 * h;
        float vw = aspect * vh;
        V3 e(eye[0], eye[1], eye[2]), la(lookat[0], lookat[1], lookat[2]), u0(up[0], up[1], up[2]);
        V3 w = normalize(e - la);
        V3 u = normalize(cross(u0, w));
        V3 v = cross(w, u);
        s->eye = e;
        s->horizontal = vw * u;
        s->vertical = vh * v;
        s->llc = e - s->horizontal / 2.0f - s->vertical / 2.0f - w;
    }
    // scene.cpp:164-206
    s->tris.resize(n);
    for (int i = 0; i < n; i++)
    {
        Tri &t = s->tris[i];
        for (int k = 0; k < 3; k++)
        {
            t.v[k] = V3(v9[i * 9 + 3 * k], v9[i * 9 + 3 * k + 1], v9[i * 9 + 3 * k + 2]);
            if (vn9)
                t.vn[k] = V3(vn9[i * 9 + 3 * k], vn9[i * 9 + 3 * k + 1], vn9[i * 9 + 3 * k + 2]);
            t.vt[k][0] = vt6 ? vt6[i * 6 + 2 * k] : 0.f, t.vt[k][1] = vt6 ? vt6[i * 6 + 2 * k + 1] : 0.f;
        }
        t.normal = normalize(cross(t.v[1] - t.v[0], t.v[2] - t.v[0]));
        V3 sum = t.v[0] + t.v[1] + t.v[2];
        t.center = V3(sum.x / 3.0f, sum.y / 3.0f, sum.z / 3.0f);
        t.mtl = mtl[i];
        t.src = i;
        Mat &m = s->mats[t.mtl];
        if (m.emissive)
        {
            t.emissive = true;
            m.area += calAera(t);
            t.area = m.area;
            m.tris.push_back(t);
        }
    }
    buildBVH(*s, 0, n - 1, leaf_num); // main.cpp:76
    // A.4 tie key for the pruned counter walk
    s->key.assign(n, 0);
    int leaves = 0;
    for (const Node &nd : s->nodes)
        leaves += nd.num > 0;
    int ord = 0;
    for (const Node &nd : s->nodes)
    {
        if (nd.num <= 0)
            continue;
        for (int k = 0; k < nd.num; k++)
        {
            const bool em = s->tris[nd.index + k].emissive;
            s->key[nd.index + k] = em ? (0x80000000u | ((uint32_t)(leaves - 1 - ord) << 3) | (uint32_t)k)
                                      : (((uint32_t)ord << 3) | (uint32_t)(7 - k));
        }
        ord++;
    }
    return s;
}

void orc_scene_destroy(orc_scene *s) { delete s; }
int32_t orc_num_nodes(orc_scene *s) { return (int32_t)s->nodes.size(); }

void orc_get_order(orc_scene *s, int32_t *perm)
{
    for (size_t i = 0; i < s->tris.size(); i++)
        perm[i] = s->tris[i].src;
}

void orc_get_derived(orc_scene *s, float *normal3, float *center3, double *cum_area, int32_t *emissive)
{
    for (size_t i = 0; i < s->tris.size(); i++)
    {
        const Tri &t = s->tris[i];
        if (normal3)
            normal3[3 * i] = t.normal.x, normal3[3 * i + 1] = t.normal.y, normal3[3 * i + 2] = t.normal.z;
        if (center3)
            center3[3 * i] = t.center.x, center3[3 * i + 1] = t.center.y, center3[3 * i + 2] = t.center.z;
        if (cum_area)
            cum_area[i] = t.area;
        if (emissive)
            emissive[i] = t.emissive;
    }
}

void orc_get_nodes(orc_scene *s, float *boxes, int32_t *links)
{
    for (size_t i = 0; i < s->nodes.size(); i++)
    {
        const Node &n = s->nodes[i];
        const float b[6] = {n.AA.x, n.AA.y, n.AA.z, n.BB.x, n.BB.y, n.BB.z};
        memcpy(boxes + 6 * i, b, sizeof b);
        const int32_t l[4] = {n.left, n.right, n.index, n.num};
        memcpy(links + 4 * i, l, sizeof l);
    }
}

void orc_get_camera(orc_scene *s, float *o)
{
    const V3 v[4] = {s->eye, s->llc, s->horizontal, s->vertical};
    for (int k = 0; k < 4; k++)
        o[3 * k] = v[k].x, o[3 * k + 1] = v[k].y, o[3 * k + 2] = v[k].z;
}

void orc_bvh_stats(orc_scene *s, int32_t *nodes, int32_t *leaves, int32_t *maxdepth)
{
    std::vector<int> depth(s->nodes.size(), 0);
    int l = 0, md = 0;
    for (size_t i = 0; i < s->nodes.size(); i++)
    {
        const Node &n = s->nodes[i];
        md = std::max(md, depth[i]);
        if (n.num > 0)
        {
            l++;
            continue;
        }
        if (n.left >= 0)
            depth[n.left] = depth[i] + 1;
        if (n.right >= 0)
            depth[n.right] = depth[i] + 1;
    }
    *nodes = (int32_t)s->nodes.size(), *leaves = l, *maxdepth = md;
}

void orc_trace(orc_scene *s, const float *rays6, int64_t n, float *t, int32_t *id, float *pn3, float *hitp3, int32_t threads)
{
    if (threads > 0)
        omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < n; i++)
    {
        const float *r = rays6 + 6 * i;
        Hit h = traceRoot(*s, V3(r[0], r[1], r[2]), V3(r[3], r[4], r[5]));
        if (t)
            t[i] = h.distance;
        if (id)
            id[i] = h.is_hit ? h.tri : -1;
        if (pn3)
            pn3[3 * i] = h.pn.x, pn3[3 * i + 1] = h.pn.y, pn3[3 * i + 2] = h.pn.z;
        if (hitp3)
            hitp3[3 * i] = h.hitpoint.x, hitp3[3 * i + 1] = h.hitpoint.y, hitp3[3 * i + 2] = h.hitpoint.z;
    }
}

void orc_trace_counts(orc_scene *s, const float *rays6, int64_t n, int32_t mode, uint64_t *box_tests, uint64_t *tri_tests,
                      int32_t *id_out, int32_t threads)
{
    if (threads > 0)
        omp_set_num_threads(threads);
    uint64_t box = 0, tri = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : box, tri)
    for (int64_t i = 0; i < n; i++)
    {
        const float *r = rays6 + 6 * i;
        Counters c;
        int id;
        if (mode == 0)
        {
            Hit h = traceRoot(*s, V3(r[0], r[1], r[2]), V3(r[3], r[4], r[5]), &c);
            id = h.is_hit ? h.tri : -1;
        }
        else
            id = tracePruned(*s, V3(r[0], r[1], r[2]), V3(r[3], r[4], r[5]), c);
        if (id_out)
            id_out[i] = id;
        box += c.box, tri += c.tri;
    }
    *box_tests = box, *tri_tests = tri;
}

void orc_render(orc_scene *s, int32_t spp, int32_t sample_begin, int32_t sample_end, int32_t max_depth, uint64_t seed,
                double *image, uint64_t *ray_counts, int32_t threads)
{
    if (threads > 0)
        omp_set_num_threads(threads);
    const int W = s->W, H = s->H;
    uint64_t c0 = 0, c1 = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : c0, c1)
    for (int pix = 0; pix < W * H; pix++)
    {
        const int i = pix / W, j = pix % W;
        uint64_t counts[2] = {0, 0};
        for (int k = sample_begin; k < sample_end; k++) // main.cpp:81-108 for one pixel
        {
            Rng rng{seed, (uint32_t)pix, (uint32_t)k};
            V3 S, d;
            primaryRay(*s, i, j, rng, S, d);
            Hit rec = traceRoot(*s, S, d);
            counts[0]++;
            V3 color;
            if (rec.is_hit)
                color = shade(*s, rec, -d, rng, 0, max_depth, counts) / (float)spp;
            image[3 * (size_t)pix + 0] += color.x;
            image[3 * (size_t)pix + 1] += color.y;
            image[3 * (size_t)pix + 2] += color.z;
        }
        c0 += counts[0], c1 += counts[1];
    }
    if (ray_counts)
        ray_counts[0] += c0, ray_counts[1] += c1;
}

void orc_render_pixels(orc_scene *s, int32_t spp, int32_t sample_begin, int32_t sample_end, int32_t max_depth,
                       uint64_t seed, const int32_t *pixels, int32_t n_pixels, double *rgb_out, int32_t threads)
{
    if (threads > 0)
        omp_set_num_threads(threads);
    const int W = s->W;
#pragma omp parallel for schedule(dynamic, 4)
    for (int q = 0; q < n_pixels; q++)
    {
        const int pix = pixels[q], i = pix / W, j = pix % W;
        for (int k = sample_begin; k < sample_end; k++) // main.cpp:81-108 for one pixel
        {
            Rng rng{seed, (uint32_t)pix, (uint32_t)k};
            V3 S, d;
            primaryRay(*s, i, j, rng, S, d);
            Hit rec = traceRoot(*s, S, d);
            V3 color;
            if (rec.is_hit)
                color = shade(*s, rec, -d, rng, 0, max_depth, nullptr) / (float)spp;
            rgb_out[3 * (size_t)q + 0] += color.x;
            rgb_out[3 * (size_t)q + 1] += color.y;
            rgb_out[3 * (size_t)q + 2] += color.z;
        }
    }
}

void orc_shade_batch(orc_scene *s, const float *rays6, const int32_t *tri_id, const float *t, int64_t n, uint64_t seed,
                     int32_t sample, int32_t max_depth, float *radiance3, int32_t threads)
{
    if (threads > 0)
        omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++)
    {
        V3 L;
        if (tri_id[i] >= 0)
        {
            const float *r = rays6 + 6 * i;
            const V3 S(r[0], r[1], r[2]), d(r[3], r[4], r[5]);
            const Tri &T = s->tris[tri_id[i]];
            Hit rec; // the record traverseBVH would have returned for this ray (bvh.cpp:219-226)
            rec.is_hit = true, rec.distance = t[i], rec.direction = d, rec.tri = tri_id[i], rec.emissive = T.emissive;
            rec.hitpoint = S + d * t[i];
            const V3 bc = findBaryCor(T, rec.hitpoint);
            rec.pn = normalize((T.vn[0] * bc.x) + (T.vn[1] * bc.y) + (T.vn[2] * bc.z));
            Rng rng{seed, (uint32_t)i, (uint32_t)sample};
            L = shade(*s, rec, -d, rng, 0, max_depth, nullptr);
        }
        radiance3[3 * i] = L.x, radiance3[3 * i + 1] = L.y, radiance3[3 * i + 2] = L.z;
    }
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox(ctr, key, out); }

double orc_uniform(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t depth, uint32_t slot)
{
    Rng r{seed, pixel, sample};
    return r.u(depth, slot);
}

void orc_primary_ray(orc_scene *s, int32_t i, int32_t j, int32_t k, uint64_t seed, float *ray6)
{
    Rng rng{seed, (uint32_t)(i * s->W + j), (uint32_t)k};
    V3 S, d;
    primaryRay(*s, i, j, rng, S, d);
    ray6[0] = S.x, ray6[1] = S.y, ray6[2] = S.z, ray6[3] = d.x, ray6[4] = d.y, ray6[5] = d.z;
}

int32_t orc_max_threads(void) { return omp_get_max_threads(); }
}
