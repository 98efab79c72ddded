// Host-side (no GPU) pieces of the path: reference-element tables, quadrature rules, the reference's
// procedural mesh generators, Hex8->Hex27 refinement, Lame conversion and the sequential greedy colouring.
// These are preprocessing steps that the reference also runs serially on the host
// (src/mesh/procedural.rs, src/mesh_convert.rs, fenris-quadrature, fenris-paradis/src/coloring.rs);
// the per-element hot loop itself lives in assemble.cu.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <unordered_map>
#include <vector>

#include "fb200_internal.h"

namespace fb200 {

bool element_info(int t, ElementInfo* out) {
    switch (t) {
        case FB200_QUAD4: *out = {4, 4, 2}; return true;
        case FB200_TET4: *out = {4, 4, 3}; return true;
        case FB200_HEX8: *out = {8, 8, 3}; return true;
        case FB200_HEX27: *out = {27, 8, 3}; return true;
        case FB200_TET10: *out = {10, 4, 3}; return true;
        case FB200_HEX20: *out = {20, 8, 3}; return true;
        default: return false;
    }
}

int geometry_type(int t) {
    if (t == FB200_HEX27 || t == FB200_HEX20) return FB200_HEX8;
    if (t == FB200_TET10) return FB200_TET4;
    return t;
}

namespace {
// Node signs of the tensor-product elements in the reference's (gmsh) ordering:
// Quad4 quadrilateral.rs:84-89, Hex8 hexahedron.rs:50-59, Hex27 hexahedron.rs:176-212.
const double kQuad[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
const double kHex[27][3] = {
    {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1},
    {0, -1, -1}, {-1, 0, -1}, {-1, -1, 0}, {1, 0, -1}, {1, -1, 0}, {0, 1, -1}, {1, 1, 0}, {-1, 1, 0},
    {0, -1, 1}, {-1, 0, 1}, {1, 0, 1}, {0, 1, 1},
    {0, 0, -1}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1},
    {0, 0, 0}};
const double kTetGrad[4][3] = {{-0.5, -0.5, -0.5}, {0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 0.5}};  // tetrahedron.rs:561-568
const int kTet10Edges[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {2, 3}, {1, 3}};             // tetrahedron.rs:236-241

// 1-D Lagrange helpers, src/element.rs:246-298
inline double lin(double a, double x) { return (1.0 + a * x) / 2.0; }
inline double dlin(double a) { return a / 2.0; }
inline double quad(double a, double x) {
    const double a2 = a * a;
    return (3.0 / 2.0 * a2 - 1.0) * (x * x) + 0.5 * a * x + 1.0 - a2;
}
inline double dquad(double a, double x) {
    const double a2 = a * a;
    return 2.0 * (3.0 / 2.0 * a2 - 1.0) * x + 0.5 * a;
}
}  // namespace

// G_ref(xi) stored node-major: g[node*d + i] = d phi_node / d xi_i.
void reference_gradients(int t, const double* xi, double* g) {
    switch (t) {
        case FB200_QUAD4:
            for (int k = 0; k < 4; ++k) {
                g[2 * k + 0] = kQuad[k][0] * (1.0 + kQuad[k][1] * xi[1]) / 4.0;
                g[2 * k + 1] = kQuad[k][1] * (1.0 + kQuad[k][0] * xi[0]) / 4.0;
            }
            break;
        case FB200_TET4:
            for (int k = 0; k < 4; ++k)
                for (int i = 0; i < 3; ++i) g[3 * k + i] = kTetGrad[k][i];
            break;
        case FB200_TET10: {  // tetrahedron.rs:198-223
            const double psi[4] = {-0.5 * xi[0] - 0.5 * xi[1] - 0.5 * xi[2] - 0.5, 0.5 * xi[0] + 0.5, 0.5 * xi[1] + 0.5,
                                   0.5 * xi[2] + 0.5};
            for (int k = 0; k < 4; ++k)
                for (int i = 0; i < 3; ++i) g[3 * k + i] = kTetGrad[k][i] * (4.0 * psi[k] - 1.0);
            for (int e = 0; e < 6; ++e) {
                const int a = kTet10Edges[e][0], b = kTet10Edges[e][1];
                for (int i = 0; i < 3; ++i) g[3 * (4 + e) + i] = kTetGrad[a][i] * (4.0 * psi[b]) + kTetGrad[b][i] * (4.0 * psi[a]);
            }
            break;
        }
        case FB200_HEX8:  // hexahedron.rs:63-83
            for (int k = 0; k < 8; ++k) {
                const double a = kHex[k][0], b = kHex[k][1], c = kHex[k][2];
                g[3 * k + 0] = dlin(a) * lin(b, xi[1]) * lin(c, xi[2]);
                g[3 * k + 1] = lin(a, xi[0]) * dlin(b) * lin(c, xi[2]);
                g[3 * k + 2] = lin(a, xi[0]) * lin(b, xi[1]) * dlin(c);
            }
            break;
        case FB200_HEX27:  // hexahedron.rs:269-315
            for (int k = 0; k < 27; ++k) {
                const double a = kHex[k][0], b = kHex[k][1], c = kHex[k][2];
                g[3 * k + 0] = dquad(a, xi[0]) * quad(b, xi[1]) * quad(c, xi[2]);
                g[3 * k + 1] = quad(a, xi[0]) * dquad(b, xi[1]) * quad(c, xi[2]);
                g[3 * k + 2] = quad(a, xi[0]) * quad(b, xi[1]) * dquad(c, xi[2]);
            }
            break;
        case FB200_HEX20:  // serendipity hexahedron, hexahedron.rs:467-545: corners phi = f g / 8, edges phi = h g / 4
            for (int k = 0; k < 20; ++k) {
                const double al = kHex[k][0], be = kHex[k][1], ga = kHex[k][2];
                const double x0 = xi[0], x1 = xi[1], x2 = xi[2];
                const double gg = (1.0 + al * x0) * (1.0 + be * x1) * (1.0 + ga * x2);
                if (k < 8) {
                    const double f = al * x0 + be * x1 + ga * x2 - 2.0, s = 1.0 / 8.0;
                    g[3 * k + 0] = s * (al * gg + f * al * (1.0 + be * x1) * (1.0 + ga * x2));
                    g[3 * k + 1] = s * (be * gg + f * be * (1.0 + al * x0) * (1.0 + ga * x2));
                    g[3 * k + 2] = s * (ga * gg + f * ga * (1.0 + al * x0) * (1.0 + be * x1));
                } else {
                    const double a2 = al * al, b2 = be * be, c2 = ga * ga, s = 1.0 / 4.0;
                    const double h = (1.0 - (1.0 - a2) * x0 * x0) * (1.0 - (1.0 - b2) * x1 * x1) * (1.0 - (1.0 - c2) * x2 * x2);
                    const double dh0 = -2.0 * (1.0 - a2) * x0 * (1.0 - (1.0 - b2) * x1 * x1) * (1.0 - (1.0 - c2) * x2 * x2);
                    const double dh1 = -2.0 * (1.0 - b2) * x1 * (1.0 - (1.0 - a2) * x0 * x0) * (1.0 - (1.0 - c2) * x2 * x2);
                    const double dh2 = -2.0 * (1.0 - c2) * x2 * (1.0 - (1.0 - a2) * x0 * x0) * (1.0 - (1.0 - b2) * x1 * x1);
                    g[3 * k + 0] = s * (dh0 * gg + h * al * (1.0 + be * x1) * (1.0 + ga * x2));
                    g[3 * k + 1] = s * (dh1 * gg + h * be * (1.0 + al * x0) * (1.0 + ga * x2));
                    g[3 * k + 2] = s * (dh2 * gg + h * ga * (1.0 + al * x0) * (1.0 + be * x1));
                }
            }
            break;
        default: break;
    }
}

// phi(xi): Quad4 quadrilateral.rs:79-91, Tet4 tetrahedron.rs:551-558, Tet10 tetrahedron.rs:179-196, Hex8 hexahedron.rs:43-60,
// Hex27 hexahedron.rs:223-268 (mass matrix / source vector / map_reference_coords)
void reference_basis(int t, const double* xi, double* phi) {
    switch (t) {
        case FB200_QUAD4:
            for (int k = 0; k < 4; ++k) phi[k] = lin(kQuad[k][0], xi[0]) * lin(kQuad[k][1], xi[1]);
            break;
        case FB200_TET4:
        case FB200_TET10: {
            const double psi[4] = {-0.5 * xi[0] - 0.5 * xi[1] - 0.5 * xi[2] - 0.5, 0.5 * xi[0] + 0.5, 0.5 * xi[1] + 0.5,
                                   0.5 * xi[2] + 0.5};
            if (t == FB200_TET4) {
                for (int k = 0; k < 4; ++k) phi[k] = psi[k];
            } else {
                for (int k = 0; k < 4; ++k) phi[k] = psi[k] * (2.0 * psi[k] - 1.0);
                for (int e = 0; e < 6; ++e) phi[4 + e] = 4.0 * psi[kTet10Edges[e][0]] * psi[kTet10Edges[e][1]];
            }
            break;
        }
        case FB200_HEX8:
            for (int k = 0; k < 8; ++k) phi[k] = lin(kHex[k][0], xi[0]) * lin(kHex[k][1], xi[1]) * lin(kHex[k][2], xi[2]);
            break;
        case FB200_HEX27:
            for (int k = 0; k < 27; ++k) phi[k] = quad(kHex[k][0], xi[0]) * quad(kHex[k][1], xi[1]) * quad(kHex[k][2], xi[2]);
            break;
        case FB200_HEX20:  // hexahedron.rs:414-465
            for (int k = 0; k < 20; ++k) {
                const double al = kHex[k][0], be = kHex[k][1], ga = kHex[k][2];
                const double gg = (1.0 + al * xi[0]) * (1.0 + be * xi[1]) * (1.0 + ga * xi[2]);
                if (k < 8)
                    phi[k] = (1.0 / 8.0) * gg * (al * xi[0] + be * xi[1] + ga * xi[2] - 2.0);
                else
                    phi[k] = (1.0 / 4.0) * (1.0 - (1.0 - al * al) * xi[0] * xi[0]) * (1.0 - (1.0 - be * be) * xi[1] * xi[1]) *
                             (1.0 - (1.0 - ga * ga) * xi[2] * xi[2]) * gg;
            }
            break;
        default: break;
    }
}

namespace {
// Hex8 basis values, hexahedron.rs:43-60 (used to place Hex27 face/centre nodes)
void hex8_basis(const double* xi, double* N) {
    for (int k = 0; k < 8; ++k) N[k] = lin(kHex[k][0], xi[0]) * lin(kHex[k][1], xi[1]) * lin(kHex[k][2], xi[2]);
}

// Gauss-Legendre on [-1,1] by Newton on the Legendre recurrence; positive roots first, then mirrored
// (fenris-quadrature/src/univariate.rs:22-36,66-117).
void legendre(int n, double x, double* p1o, double* p2o) {
    double p1 = 1.0, p2 = 0.0;
    for (int m = 1; m <= n; ++m) {
        const double mf = (double)m;
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * mf - 1.0) * x * p2 - (mf - 1.0) * p3) / mf;
    }
    *p1o = p1;
    *p2o = p2;
}
void gauss(int n, std::vector<double>& w, std::vector<double>& x) {
    const int m = (n + 1) / 2;
    w.clear();
    x.clear();
    for (int i = 0; i < m; ++i) {
        double xi = std::cos(M_PI * ((double)i + 0.75) / ((double)n + 0.5));
        double p1, p2;
        legendre(n, xi, &p1, &p2);
        double p = p1, dp = (double)n * (xi * p1 - p2) / (xi * xi - 1.0);
        for (;;) {
            const double dx = -p / dp;
            xi += dx;
            legendre(n, xi, &p1, &p2);
            p = p1;
            dp = (double)n * (xi * p1 - p2) / (xi * xi - 1.0);
            if (std::fabs(dx) <= 1e-15) break;
        }
        x.push_back(xi);
        w.push_back(2.0 / ((1.0 - xi * xi) * dp * dp));
    }
    for (int i = m; i < n; ++i) {
        const int mirror = n - i - 1;
        x.push_back(-x[mirror]);
        w.push_back(w[mirror]);
    }
}
}  // namespace

}  // namespace fb200

using namespace fb200;

extern "C" {

void fb200_lame_from_young_poisson(double young, double poisson, double* mu, double* lambda) {
    const double m = 0.5 * young / (1.0 + poisson);
    *mu = m;
    *lambda = 2.0 * m * poisson / (1.0 - 2.0 * poisson);
}

fb200_status fb200_canonical_quadrature(int32_t element_type, int32_t* num_points, double* weights, double* points) {
    if (!num_points) return FB200_ERR_SHAPE;
    std::vector<double> w1, x1;
    switch (element_type) {
        case FB200_QUAD4: {  // tensor.rs:13-32, x outer / y inner
            gauss(2, w1, x1);
            *num_points = 4;
            if (!weights || !points) return FB200_OK;
            int k = 0;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b, ++k) {
                    weights[k] = w1[a] * w1[b];
                    points[2 * k] = x1[a];
                    points[2 * k + 1] = x1[b];
                }
            return FB200_OK;
        }
        case FB200_HEX8:
        case FB200_HEX20:
        case FB200_HEX27: {  // tensor.rs:36-58, x outer / z inner; canonical.rs:102-112: Gauss 2^3 for Hex8, 3^3 for Hex20 / Hex27
            const int n = element_type == FB200_HEX8 ? 2 : 3;
            gauss(n, w1, x1);
            *num_points = n * n * n;
            if (!weights || !points) return FB200_OK;
            int k = 0;
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b)
                    for (int c = 0; c < n; ++c, ++k) {
                        weights[k] = w1[a] * w1[b] * w1[c];
                        points[3 * k] = x1[a];
                        points[3 * k + 1] = x1[b];
                        points[3 * k + 2] = x1[c];
                    }
            return FB200_OK;
        }
        case FB200_TET4:  // rules/polyquad/expanded/tet/1-1.txt
            *num_points = 1;
            if (!weights || !points) return FB200_OK;
            weights[0] = 1.3333333333333333333333333333333333333;
            points[0] = points[1] = points[2] = -0.5;
            return FB200_OK;
        case FB200_TET10: {  // rules/polyquad/expanded/tet/2-4.txt
            *num_points = 4;
            if (!weights || !points) return FB200_OK;
            const double a = -0.72360679774997896964091736687312762354, b = 0.17082039324993690892275210061938287063;
            const double pts[4][3] = {{a, a, b}, {a, b, a}, {b, a, a}, {a, a, a}};
            for (int k = 0; k < 4; ++k) {
                weights[k] = 0.33333333333333333333333333333333333333;
                for (int i = 0; i < 3; ++i) points[3 * k + i] = pts[k][i];
            }
            return FB200_OK;
        }
        default: return FB200_ERR_UNSUPPORTED;
    }
}

// ---------------------------------------------------------------- generators (src/mesh/procedural.rs)
fb200_status fb200_gen_hex_mesh(uint64_t cx, uint64_t cy, uint64_t cz, double h, uint64_t* nv, uint64_t* ne, double* v,
                                uint64_t* conn) {
    const uint64_t vx = cx + 1, vy = cy + 1, vz = cz + 1;
    if (cx == 0 || cy == 0 || cz == 0) {
        if (nv) *nv = 0;
        if (ne) *ne = 0;
        return FB200_OK;
    }
    if (nv) *nv = vx * vy * vz;
    if (ne) *ne = cx * cy * cz;
    if (!v || !conn) return FB200_OK;
    uint64_t p = 0;
    for (uint64_t k = 0; k < vz; ++k)
        for (uint64_t j = 0; j < vy; ++j)
            for (uint64_t i = 0; i < vx; ++i) {
                v[p++] = (double)i * h;
                v[p++] = (double)j * h;
                v[p++] = (double)k * h;
            }
    auto idx = [&](uint64_t i, uint64_t j, uint64_t k) { return (vx * vy) * k + vx * j + i; };
    p = 0;
    for (uint64_t k = 0; k < cz; ++k)
        for (uint64_t j = 0; j < cy; ++j)
            for (uint64_t i = 0; i < cx; ++i) {
                conn[p++] = idx(i, j, k);
                conn[p++] = idx(i + 1, j, k);
                conn[p++] = idx(i + 1, j + 1, k);
                conn[p++] = idx(i, j + 1, k);
                conn[p++] = idx(i, j, k + 1);
                conn[p++] = idx(i + 1, j, k + 1);
                conn[p++] = idx(i + 1, j + 1, k + 1);
                conn[p++] = idx(i, j + 1, k + 1);
            }
    return FB200_OK;
}

fb200_status fb200_gen_quad_mesh(uint64_t cx, uint64_t cy, double h, uint64_t* nv, uint64_t* ne, double* v, uint64_t* conn) {
    if (cx == 0 || cy == 0) {
        if (nv) *nv = 0;
        if (ne) *ne = 0;
        return FB200_OK;
    }
    if (nv) *nv = (cx + 1) * (cy + 1);
    if (ne) *ne = cx * cy;
    if (!v || !conn) return FB200_OK;
    uint64_t p = 0;
    // top_left = (0, 1): rows run downwards (procedural.rs:69-76)
    for (uint64_t j = 0; j <= cy; ++j)
        for (uint64_t i = 0; i <= cx; ++i) {
            v[p++] = 0.0 + (double)i * h;
            v[p++] = 1.0 + (-(double)j) * h;
        }
    auto idx = [&](uint64_t i, uint64_t j) { return (cx + 1) * j + i; };
    p = 0;
    for (uint64_t j = 0; j < cy; ++j)
        for (uint64_t i = 0; i < cx; ++i) {
            conn[p++] = idx(i, j + 1);
            conn[p++] = idx(i + 1, j + 1);
            conn[p++] = idx(i + 1, j);
            conn[p++] = idx(i, j);
        }
    return FB200_OK;
}

fb200_status fb200_gen_tet_mesh(uint64_t cx, uint64_t cy, uint64_t cz, double h, uint64_t* nv, uint64_t* ne, double* v,
                                uint64_t* conn) {
    if (cx == 0 || cy == 0 || cz == 0) {
        if (nv) *nv = 0;
        if (ne) *ne = 0;
        return FB200_OK;
    }
    const uint64_t vx = cx + 1, vy = cy + 1, vz = cz + 1;
    const uint64_t centre0 = vx * vy * vz;
    // 4 tets per interior face pair + 2 per boundary face: 12 per cell on a box
    const uint64_t ntet = 4 * ((cx - 1) * cy * cz + cx * (cy - 1) * cz + cx * cy * (cz - 1)) + 2 * 2 * (cy * cz + cx * cz + cx * cy);
    if (nv) *nv = centre0 + cx * cy * cz;
    if (ne) *ne = ntet;
    if (!v || !conn) return FB200_OK;
    uint64_t p = 0;
    for (uint64_t k = 0; k < vz; ++k)
        for (uint64_t j = 0; j < vy; ++j)
            for (uint64_t i = 0; i < vx; ++i) {
                v[p++] = h * (double)i;
                v[p++] = h * (double)j;
                v[p++] = h * (double)k;
            }
    for (uint64_t k = 0; k < cz; ++k)
        for (uint64_t j = 0; j < cy; ++j)
            for (uint64_t i = 0; i < cx; ++i) {
                v[p++] = h * (0.5 + (double)i);
                v[p++] = h * (0.5 + (double)j);
                v[p++] = h * (0.5 + (double)k);
            }
    // face of cell (i,j,k) shared with its +axis neighbour, as a vertex cycle (procedural.rs:331-335)
    static const int face[3][4][3] = {{{1, 0, 1}, {1, 1, 1}, {1, 1, 0}, {1, 0, 0}},
                                      {{0, 1, 0}, {1, 1, 0}, {1, 1, 1}, {0, 1, 1}},
                                      {{0, 1, 1}, {1, 1, 1}, {1, 0, 1}, {0, 0, 1}}};
    auto vid = [&](int64_t i, int64_t j, int64_t k) { return (vx * vy) * (uint64_t)k + vx * (uint64_t)j + (uint64_t)i; };
    auto cid = [&](uint64_t i, uint64_t j, uint64_t k) { return (cx * cy) * k + cx * j + i + centre0; };
    const uint64_t nc[3] = {cx, cy, cz};
    uint64_t t = 0;
    auto emit = [&](uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
        conn[t++] = a;
        conn[t++] = b;
        conn[t++] = c;
        conn[t++] = d;
    };
    for (uint64_t k = 0; k < cz; ++k)
        for (uint64_t j = 0; j < cy; ++j)
            for (uint64_t i = 0; i < cx; ++i) {
                const uint64_t cell[3] = {i, j, k};
                for (int axis = 0; axis < 3; ++axis) {
                    if (cell[axis] + 1 < nc[axis]) {  // octahedron between two cell centres -> 4 tets
                        uint64_t f[4];
                        for (int m = 0; m < 4; ++m) f[m] = vid(i + face[axis][m][0], j + face[axis][m][1], k + face[axis][m][2]);
                        uint64_t nb[3] = {i, j, k};
                        nb[axis] += 1;
                        const uint64_t c1 = cid(i, j, k), c2 = cid(nb[0], nb[1], nb[2]);
                        for (int m = 0; m < 4; ++m) emit(c1, c2, f[(m + 1) & 3], f[m]);
                    }
                    for (int side = 0; side < 2; ++side) {  // boundary pyramids -> 2 tets, alternating diagonal
                        const bool low = side == 0;
                        if ((low && cell[axis] != 0) || (!low && cell[axis] + 1 != nc[axis])) continue;
                        int64_t fv[4][3];
                        for (int m = 0; m < 4; ++m) {
                            const int src = low ? 3 - m : m;  // reversed orientation on the low side
                            fv[m][0] = (int64_t)i + face[axis][src][0];
                            fv[m][1] = (int64_t)j + face[axis][src][1];
                            fv[m][2] = (int64_t)k + face[axis][src][2];
                            if (low) fv[m][axis] -= 1;
                        }
                        const uint64_t a = vid(fv[0][0], fv[0][1], fv[0][2]), b = vid(fv[1][0], fv[1][1], fv[1][2]);
                        const uint64_t c = vid(fv[2][0], fv[2][1], fv[2][2]), d = vid(fv[3][0], fv[3][1], fv[3][2]);
                        const uint64_t ctr = cid(i, j, k);
                        if ((i + j + k) % 2 == 0) {
                            emit(a, b, c, ctr);
                            emit(a, c, d, ctr);
                        } else {
                            emit(a, b, d, ctr);
                            emit(b, c, d, ctr);
                        }
                    }
                }
            }
    return t == 4 * ntet ? FB200_OK : FB200_ERR_SHAPE;
}

// ---------------------------------------------------------------- Hex8 -> Hex27 (src/mesh_convert.rs:85-166,227-330)
static fb200_status hex_refine(int nodes_out, uint64_t nv, const double* v, uint64_t ne, const uint64_t* hex8, uint64_t* nv27, double* v27,
                               uint64_t* hex27);
fb200_status fb200_hex27_from_hex8(uint64_t nv, const double* v, uint64_t ne, const uint64_t* hex8, uint64_t* nv27, double* v27,
                                   uint64_t* hex27) {
    return hex_refine(27, nv, v, ne, hex8, nv27, v27, hex27);
}
// Hex20: the vertices and the 12 edge midpoints only (src/mesh_convert.rs:168-217)
fb200_status fb200_hex20_from_hex8(uint64_t nv, const double* v, uint64_t ne, const uint64_t* hex8, uint64_t* nv20, double* v20,
                                   uint64_t* hex20) {
    return hex_refine(20, nv, v, ne, hex8, nv20, v20, hex20);
}
static fb200_status hex_refine(int nodes_out, uint64_t nv, const double* v, uint64_t ne, const uint64_t* hex8, uint64_t* nv27, double* v27,
                               uint64_t* hex27) {
    static const int edges[12][2] = {{0, 1}, {0, 3}, {0, 4}, {1, 2}, {1, 5}, {2, 3}, {2, 6}, {3, 7}, {4, 5}, {4, 7}, {5, 6}, {6, 7}};
    static const int faces[6][4] = {{0, 1, 2, 3}, {0, 1, 4, 5}, {0, 3, 4, 7}, {1, 2, 5, 6}, {2, 3, 6, 7}, {4, 5, 6, 7}};
    static const double face_ref[6][3] = {{0, 0, -1}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    struct Key {
        uint64_t k[8];
        bool operator<(const Key& o) const { return std::lexicographical_compare(k, k + 8, o.k, o.k + 8); }
    };
    std::map<Key, uint64_t> label;  // sorted parent set -> new vertex id, first-seen order
    std::vector<double> out;
    const bool write = v27 != nullptr && hex27 != nullptr;
    for (uint64_t e = 0; e < ne; ++e) {
        const uint64_t* g = hex8 + 8 * e;
        double X[8][3];
        for (int a = 0; a < 8; ++a) {
            if (g[a] >= nv) return FB200_ERR_INDEX_OOB;
            for (int i = 0; i < 3; ++i) X[a][i] = v[3 * g[a] + i];
        }
        for (int l = 0; l < nodes_out; ++l) {
            Key key;
            std::fill(key.k, key.k + 8, ~0ull);
            double pos[3] = {0, 0, 0};
            if (l < 8) {
                key.k[0] = g[l];
                for (int i = 0; i < 3; ++i) pos[i] = X[l][i];
            } else if (l < 20) {
                const int a = edges[l - 8][0], b = edges[l - 8][1];
                key.k[0] = g[a];
                key.k[1] = g[b];
                for (int i = 0; i < 3; ++i) pos[i] = 0.5 * X[b][i] + 0.5 * X[a][i];  // lerp(a, b, 0.5)
            } else {
                double Nb[8];
                const double centre[3] = {0, 0, 0};
                hex8_basis(l < 26 ? face_ref[l - 20] : centre, Nb);
                for (int a = 0; a < 8; ++a)
                    for (int i = 0; i < 3; ++i) pos[i] += X[a][i] * Nb[a];
                if (l < 26)
                    for (int m = 0; m < 4; ++m) key.k[m] = g[faces[l - 20][m]];
                else
                    for (int m = 0; m < 8; ++m) key.k[m] = g[m];
            }
            std::sort(key.k, key.k + 8);
            auto it = label.find(key);
            uint64_t id;
            if (it == label.end()) {
                id = label.size();
                label.emplace(key, id);
                if (write) {
                    v27[3 * id] = pos[0];
                    v27[3 * id + 1] = pos[1];
                    v27[3 * id + 2] = pos[2];
                }
            } else {
                id = it->second;
            }
            if (write) hex27[(uint64_t)nodes_out * e + l] = id;
        }
    }
    if (nv27) *nv27 = label.size();
    return FB200_OK;
}

// ---------------------------------------------------------------- Tet4 -> Tet10 (src/mesh_convert.rs:42-83, 227-330, 444-452)
// The 4 vertices, then the midpoints of the edges (0,1) (1,2) (0,2) (0,3) (2,3) (1,3); global labels in first-seen order keyed by the
// sorted parent set, midpoints = lerp(a, b, 0.5) = 0.5 b + 0.5 a exactly as nalgebra evaluates it.
fb200_status fb200_tet10_from_tet4(uint64_t nv, const double* v, uint64_t ne, const uint64_t* tet4, uint64_t* nv10, double* v10,
                                   uint64_t* tet10) {
    static const int edges[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {2, 3}, {1, 3}};
    if (!v || !tet4) return FB200_ERR_SHAPE;
    std::map<std::pair<uint64_t, uint64_t>, uint64_t> label;  // (min parent, max parent | ~0 for a vertex) -> new id
    const bool write = v10 != nullptr && tet10 != nullptr;
    for (uint64_t e = 0; e < ne; ++e) {
        const uint64_t* g = tet4 + 4 * e;
        for (int a = 0; a < 4; ++a)
            if (g[a] >= nv) return FB200_ERR_INDEX_OOB;
        for (int l = 0; l < 10; ++l) {
            std::pair<uint64_t, uint64_t> key;
            double pos[3];
            if (l < 4) {
                key = {g[l], ~0ull};
                for (int i = 0; i < 3; ++i) pos[i] = v[3 * g[l] + i];
            } else {
                const uint64_t a = g[edges[l - 4][0]], b = g[edges[l - 4][1]];
                key = {std::min(a, b), std::max(a, b)};
                for (int i = 0; i < 3; ++i) pos[i] = 0.5 * v[3 * b + i] + 0.5 * v[3 * a + i];
            }
            auto it = label.find(key);
            uint64_t id;
            if (it == label.end()) {
                id = label.size();
                label.emplace(key, id);
                if (write) {
                    v10[3 * id] = pos[0];
                    v10[3 * id + 1] = pos[1];
                    v10[3 * id + 2] = pos[2];
                }
            } else {
                id = it->second;
            }
            if (write) tet10[10 * e + l] = id;
        }
    }
    if (nv10) *nv10 = label.size();
    return FB200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- greedy colouring (fenris-paradis/src/coloring.rs:6-70)
namespace fb200 {
void greedy_coloring(uint64_t E, uint64_t N, const std::vector<int64_t>& off, const std::vector<int32_t>& nodes,
                     std::vector<uint64_t>& color_off, std::vector<uint64_t>& color_elems) {
    std::vector<int32_t> last(N, -1);
    std::vector<uint64_t> cur(E), post;
    for (uint64_t e = 0; e < E; ++e) cur[e] = e;
    color_off.assign(1, 0);
    color_elems.clear();
    color_elems.reserve(E);
    int32_t c = 0;
    while (!cur.empty()) {
        post.clear();
        for (uint64_t e : cur) {
            bool blocked = false;
            for (int64_t k = off[e]; k < off[e + 1]; ++k)
                if (last[nodes[k]] == c) {
                    blocked = true;
                    break;
                }
            if (blocked) {
                post.push_back(e);
            } else {
                for (int64_t k = off[e]; k < off[e + 1]; ++k) last[nodes[k]] = c;
                color_elems.push_back(e);
            }
        }
        color_off.push_back(color_elems.size());
        cur.swap(post);
        ++c;
    }
}
}  // namespace fb200
