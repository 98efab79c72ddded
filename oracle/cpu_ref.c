/*
 * oracle/cpu_ref.c - C restatement of the reference's CPU assembly path.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE, NOT THE PRODUCT.  Only tests/,
 * __graft_entry__.smoke() and bench.py (cpu_baseline leg, --impl reference) may
 * load it.  The product library (fenris_b200/csrc) never links or calls this.
 *
 * What it restates (reference = InteractiveComputerGraphics/fenris @ 7181b15;
 * the Rust reference cannot be compiled in this environment: no rustc/cargo):
 *   - CsrAssembler::assemble_pattern            src/assembly/global.rs:65-120
 *   - CsrAssembler::assemble_into_csr           src/assembly/global.rs:133-182
 *   - CsrParAssembler::assemble_into_csr        src/assembly/global.rs:314-376  (rayon -> OpenMP)
 *   - add_element_row_to_csr_row                src/assembly/global.rs:504-537
 *   - color_nodes / sequential_greedy_coloring  src/assembly/global.rs:540-551, fenris-paradis/src/coloring.rs:6-70
 *   - assemble_element_elliptic_matrix          src/assembly/local/elliptic.rs:361-439
 *   - EllipticContraction default block loop    src/assembly/operators.rs:146-189, fenris-solid/src/lib.rs:349-392
 *   - LaplaceOperator::contract                 src/assembly/operators/laplace.rs:60-68
 *   - LinearElasticMaterial contraction         fenris-solid/src/materials.rs:108-122
 *   - clone_upper_to_lower                      src/util.rs:38-50
 *   - element tables                            src/element.rs:246-298, element/{hexahedron,tetrahedron,quadrilateral}.rs
 *   - mesh generators                           src/mesh/procedural.rs:46-93,216-277,286-403
 * nalgebra 0.32.1 (not in the reference tree) 2x2/3x3 determinant/inverse closed forms are restated
 * from its published source (src/linalg/determinant.rs, inverse.rs).
 *
 * Pinned by tests/test_cpu_ref.py against oracle/fenris_oracle.py (which is pinned to the
 * reference's golden vectors) - bitwise for patterns/meshes/colours, <=1e-14 rel. Frobenius for values.
 *
 * Like the reference, the element kernel rebuilds the element and re-evaluates the basis gradients
 * for the Jacobian and again for the basis (space_impl.rs:95-128): it is a faithful CPU baseline,
 * not a tuned one.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OREF_QUAD4 1
#define OREF_TET4 2
#define OREF_HEX8 3
#define OREF_HEX27 4
#define OREF_TET10 5

#define OREF_LAPLACE 1
#define OREF_LINEAR_ELASTIC 2

#define OREF_OK 0
#define OREF_ERR_SINGULAR 1
#define OREF_ERR_COLUMN 2
#define OREF_ERR_ARG 3

#define MAXN 27
#define MAXD 3

typedef uint64_t u64;

/* Threads the CPU path can use: the processors available to the process, NOT omp_get_max_threads() - torchrun exports
 * OMP_NUM_THREADS=1 to every rank, which would silently time the multi-rank reference arm of bench.py on one thread. */
int oref_max_threads(void) {
#ifdef _OPENMP
    int n = omp_get_num_procs();
    return n > 0 ? n : 1;
#else
    return 1;
#endif
}

static int elem_nodes(int t) {
    switch (t) { case OREF_QUAD4: return 4; case OREF_TET4: return 4; case OREF_HEX8: return 8;
                 case OREF_HEX27: return 27; case OREF_TET10: return 10; default: return 0; }
}
static int elem_geom_nodes(int t) {
    switch (t) { case OREF_QUAD4: return 4; case OREF_TET4: return 4; case OREF_HEX8: return 8;
                 case OREF_HEX27: return 8; case OREF_TET10: return 4; default: return 0; }
}
static int elem_dim(int t) { return t == OREF_QUAD4 ? 2 : 3; }
int oref_elem_nodes(int t) { return elem_nodes(t); }
int oref_elem_dim(int t) { return elem_dim(t); }

/* ---- 1-D helpers, src/element.rs:246-298 ---- */
static double phi_lin(double a, double x) { return (1.0 + a * x) / 2.0; }
static double dphi_lin(double a) { return a / 2.0; }
static double phi_quad(double a, double x) { double a2 = a * a; return (3.0 / 2.0 * a2 - 1.0) * (x * x) + 0.5 * a * x + 1.0 - a2; }
static double dphi_quad(double a, double x) { double a2 = a * a; return 2.0 * (3.0 / 2.0 * a2 - 1.0) * x + 0.5 * a; }

static const double QUAD4_N[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
static const double HEX27_N[27][3] = {
    {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1},
    {0, -1, -1}, {-1, 0, -1}, {-1, -1, 0}, {1, 0, -1}, {1, -1, 0}, {0, 1, -1}, {1, 1, 0}, {-1, 1, 0},
    {0, -1, 1}, {-1, 0, 1}, {1, 0, 1}, {0, 1, 1},
    {0, 0, -1}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1},
    {0, 0, 0}};
static const double TET4_G[4][3] = {{-0.5, -0.5, -0.5}, {0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 0.5}};
static const int TET10_E[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {2, 3}, {1, 3}};

/* reference gradients, column-major g[d*node + i] (column = node) */
static void ref_gradients(int t, const double* xi, double* g) {
    int k;
    switch (t) {
    case OREF_QUAD4: /* quadrilateral.rs:94-107 */
        for (k = 0; k < 4; ++k) {
            double a = QUAD4_N[k][0], b = QUAD4_N[k][1];
            g[2 * k + 0] = a * (1.0 + b * xi[1]) / 4.0;
            g[2 * k + 1] = b * (1.0 + a * xi[0]) / 4.0;
        }
        break;
    case OREF_TET4: /* tetrahedron.rs:561-568 */
        for (k = 0; k < 4; ++k) { g[3 * k] = TET4_G[k][0]; g[3 * k + 1] = TET4_G[k][1]; g[3 * k + 2] = TET4_G[k][2]; }
        break;
    case OREF_TET10: { /* tetrahedron.rs:198-223 */
        double psi[4];
        int i, e;
        psi[0] = -0.5 * xi[0] - 0.5 * xi[1] - 0.5 * xi[2] - 0.5;
        psi[1] = 0.5 * xi[0] + 0.5; psi[2] = 0.5 * xi[1] + 0.5; psi[3] = 0.5 * xi[2] + 0.5;
        for (k = 0; k < 4; ++k) for (i = 0; i < 3; ++i) g[3 * k + i] = TET4_G[k][i] * (4.0 * psi[k] - 1.0);
        for (e = 0; e < 6; ++e) {
            int a = TET10_E[e][0], b = TET10_E[e][1];
            for (i = 0; i < 3; ++i) g[3 * (4 + e) + i] = TET4_G[a][i] * (4.0 * psi[b]) + TET4_G[b][i] * (4.0 * psi[a]);
        }
        break;
    }
    case OREF_HEX8: /* hexahedron.rs:63-83 */
        for (k = 0; k < 8; ++k) {
            double a = HEX27_N[k][0], b = HEX27_N[k][1], c = HEX27_N[k][2];
            g[3 * k + 0] = dphi_lin(a) * phi_lin(b, xi[1]) * phi_lin(c, xi[2]);
            g[3 * k + 1] = phi_lin(a, xi[0]) * dphi_lin(b) * phi_lin(c, xi[2]);
            g[3 * k + 2] = phi_lin(a, xi[0]) * phi_lin(b, xi[1]) * dphi_lin(c);
        }
        break;
    case OREF_HEX27: /* hexahedron.rs:269-315 */
        for (k = 0; k < 27; ++k) {
            double a = HEX27_N[k][0], b = HEX27_N[k][1], c = HEX27_N[k][2];
            g[3 * k + 0] = dphi_quad(a, xi[0]) * phi_quad(b, xi[1]) * phi_quad(c, xi[2]);
            g[3 * k + 1] = phi_quad(a, xi[0]) * dphi_quad(b, xi[1]) * phi_quad(c, xi[2]);
            g[3 * k + 2] = phi_quad(a, xi[0]) * phi_quad(b, xi[1]) * dphi_quad(c, xi[2]);
        }
        break;
    }
}

static int geom_type(int t) { return t == OREF_HEX27 ? OREF_HEX8 : (t == OREF_TET10 ? OREF_TET4 : t); }

/* nalgebra closed forms; m row-major d x d */
static double det_small(int d, const double* m) {
    if (d == 2) return m[0] * m[3] - m[2] * m[1];
    {
        double m11 = m[0], m12 = m[1], m13 = m[2], m21 = m[3], m22 = m[4], m23 = m[5], m31 = m[6], m32 = m[7], m33 = m[8];
        double minor_m12_m23 = m22 * m33 - m32 * m23;
        double minor_m11_m23 = m21 * m33 - m31 * m23;
        double minor_m11_m22 = m21 * m32 - m31 * m22;
        return m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    }
}
static int inv_small(int d, const double* m, double det, double* inv) {
    if (det == 0.0) return 0;
    if (d == 2) {
        inv[0] = m[3] / det; inv[1] = -m[1] / det; inv[2] = -m[2] / det; inv[3] = m[0] / det;
        return 1;
    }
    {
        double m11 = m[0], m12 = m[1], m13 = m[2], m21 = m[3], m22 = m[4], m23 = m[5], m31 = m[6], m32 = m[7], m33 = m[8];
        inv[0] = (m22 * m33 - m32 * m23) / det;
        inv[1] = (m13 * m32 - m33 * m12) / det;
        inv[2] = (m12 * m23 - m22 * m13) / det;
        inv[3] = -(m21 * m33 - m31 * m23) / det;
        inv[4] = (m11 * m33 - m31 * m13) / det;
        inv[5] = (m13 * m21 - m23 * m11) / det;
        inv[6] = (m21 * m32 - m31 * m22) / det;
        inv[7] = (m12 * m31 - m32 * m11) / det;
        inv[8] = (m11 * m22 - m21 * m12) / det;
        return 1;
    }
}

typedef struct {
    int elem_type, op, n, ng, d, s, q;
    const double* weights; /* q */
    const double* points;  /* q x d */
    const double* params;  /* q x 2 (mu, lambda) or NULL */
} oref_problem;

/* assemble_element_elliptic_matrix (elliptic.rs:361-439).  K column-major (s n)^2.
 * X: n x d coordinates of the element's nodes (local order). */
static int element_matrix(const oref_problem* p, const double* X, double* K) {
    const int n = p->n, ng = p->ng, d = p->d, s = p->s, sn = s * n;
    double gg[MAXD * MAXN], g[MAXD * MAXN], J[9], Jinv[9];
    int q, a, i, j, In, Jn;
    memset(K, 0, sizeof(double) * (size_t)sn * sn);
    for (q = 0; q < p->q; ++q) {
        const double* xi = p->points + (size_t)q * d;
        double w = p->weights[q], det, scale;
        /* element.reference_jacobian(point): J = X * G^T over the geometry nodes */
        ref_gradients(geom_type(p->elem_type), xi, gg);
        for (i = 0; i < d; ++i)
            for (j = 0; j < d; ++j) {
                double acc = 0.0;
                for (a = 0; a < ng; ++a) acc += X[(size_t)a * d + i] * gg[d * a + j];
                J[i * d + j] = acc;
            }
        det = det_small(d, J);
        if (!inv_small(d, J, det, Jinv)) return OREF_ERR_SINGULAR;
        /* element.populate_basis_gradients (second evaluation, like the reference) */
        ref_gradients(p->elem_type, xi, g);
        /* phi_grad <- J^{-T} phi_grad, column by column (elliptic.rs:415-418) */
        for (a = 0; a < n; ++a) {
            double t[MAXD];
            for (i = 0; i < d; ++i) {
                double acc = 0.0;
                for (j = 0; j < d; ++j) acc += Jinv[j * d + i] * g[d * a + j];
                t[i] = acc;
            }
            for (i = 0; i < d; ++i) g[d * a + i] = t[i];
        }
        scale = w * fabs(det);
        for (Jn = 0; Jn < n; ++Jn) {
            for (In = 0; In <= Jn; ++In) {
                const double* av = g + d * In;
                const double* bv = g + d * Jn;
                double dot = 0.0;
                for (i = 0; i < d; ++i) dot += av[i] * bv[i];
                if (p->op == OREF_LAPLACE) {
                    K[(size_t)Jn * sn + In] += dot * scale;
                } else {
                    double mu = p->params[2 * q], lam = p->params[2 * q + 1];
                    for (j = 0; j < d; ++j)
                        for (i = 0; i < d; ++i) {
                            double c = ((i == j ? dot : 0.0) + bv[i] * av[j]) * mu + (av[i] * bv[j]) * lam;
                            K[(size_t)(s * Jn + j) * sn + (s * In + i)] += scale * c;
                        }
                }
            }
        }
    }
    /* clone_upper_to_lower (util.rs:38-50) */
    for (j = 0; j < sn; ++j)
        for (i = j + 1; i < sn; ++i) K[(size_t)j * sn + i] = K[(size_t)i * sn + j];
    return OREF_OK;
}

int oref_element_matrix(int elem_type, int op, int q, const double* weights, const double* points, const double* params,
                        const double* X, double* K) {
    oref_problem p;
    p.elem_type = elem_type; p.op = op; p.n = elem_nodes(elem_type); p.ng = elem_geom_nodes(elem_type);
    p.d = elem_dim(elem_type); p.s = op == OREF_LAPLACE ? 1 : p.d; p.q = q;
    p.weights = weights; p.points = points; p.params = params;
    if (p.n == 0) return OREF_ERR_ARG;
    return element_matrix(&p, X, K);
}

/* sort permutation of local nodes by global id (global.rs:155-159; sort_unstable_by_key -> insertion sort, tiny n) */
static void sort_perm(int n, const u64* nodes, int* perm) {
    int i, j;
    for (i = 0; i < n; ++i) perm[i] = i;
    for (i = 1; i < n; ++i) {
        int p = perm[i];
        for (j = i; j > 0 && nodes[perm[j - 1]] > nodes[p]; --j) perm[j] = perm[j - 1];
        perm[j] = p;
    }
}

/* add_element_row_to_csr_row (global.rs:504-537) */
static int add_row(double* rv, const u64* rc, u64 rlen, const u64* nodes, const int* perm, int n, int dim,
                   const double* K, int sn, int lrow) {
    u64 cur = 0;
    int k, i;
    for (k = 0; k < n; ++k) {
        int nl = perm[k];
        u64 ng = nodes[nl];
        for (i = 0; i < dim; ++i) {
            int lcol = dim * nl + i;
            u64 gcol = (u64)dim * ng + (u64)i;
            while (cur < rlen && rc[cur] != gcol) ++cur;
            if (cur >= rlen) return OREF_ERR_COLUMN;
            rv[cur] += K[(size_t)lcol * sn + lrow]; /* local_row[local_col], K column-major */
            ++cur;
        }
    }
    return OREF_OK;
}

static int scatter_element(const oref_problem* p, const u64* row_offsets, const u64* col_indices, double* values,
                           const u64* nodes, const double* K) {
    int perm[MAXN], ln, i, st;
    const int n = p->n, s = p->s, sn = s * n;
    sort_perm(n, nodes, perm);
    for (ln = 0; ln < n; ++ln)
        for (i = 0; i < s; ++i) {
            u64 grow = (u64)s * nodes[ln] + (u64)i;
            u64 b = row_offsets[grow], e = row_offsets[grow + 1];
            st = add_row(values + b, col_indices + b, e - b, nodes, perm, n, s, K, sn, s * ln + i);
            if (st) return st;
        }
    return OREF_OK;
}

static void gather_coords(const oref_problem* p, const double* vertices, const u64* nodes, double* X) {
    int a, i;
    for (a = 0; a < p->n; ++a)
        for (i = 0; i < p->d; ++i) X[a * p->d + i] = vertices[nodes[a] * (u64)p->d + (u64)i];
}

static int init_problem(oref_problem* p, int elem_type, int op, int q, const double* w, const double* pts, const double* params) {
    p->elem_type = elem_type; p->op = op; p->n = elem_nodes(elem_type); p->ng = elem_geom_nodes(elem_type);
    p->d = elem_dim(elem_type); p->s = op == OREF_LAPLACE ? 1 : p->d; p->q = q;
    p->weights = w; p->points = pts; p->params = params;
    if (p->n == 0 || (op != OREF_LAPLACE && op != OREF_LINEAR_ELASTIC)) return OREF_ERR_ARG;
    if (op == OREF_LINEAR_ELASTIC && !params) return OREF_ERR_ARG;
    return OREF_OK;
}

/* CsrAssembler::assemble_into_csr (global.rs:133-182): serial, accumulating into `values`. */
int oref_assemble_serial(int elem_type, int op, int q, const double* w, const double* pts, const double* params,
                         const double* vertices, u64 num_elements, const u64* conn,
                         const u64* row_offsets, const u64* col_indices, double* values, int64_t* bad_element) {
    oref_problem p;
    double X[MAXN * MAXD];
    double* K;
    u64 e;
    int st = init_problem(&p, elem_type, op, q, w, pts, params);
    if (st) return st;
    K = (double*)malloc(sizeof(double) * (size_t)(p.s * p.n) * (p.s * p.n));
    for (e = 0; e < num_elements; ++e) {
        const u64* nodes = conn + e * (u64)p.n;
        gather_coords(&p, vertices, nodes, X);
        st = element_matrix(&p, X, K);
        if (!st) st = scatter_element(&p, row_offsets, col_indices, values, nodes, K);
        if (st) { if (bad_element) *bad_element = (int64_t)e; break; }
    }
    free(K);
    return st;
}

/* CsrParAssembler::assemble_into_csr (global.rs:314-376): colours sequential, elements of a colour in
 * parallel (rayon work stealing -> OpenMP dynamic schedule), thread-local K_e workspace. */
int oref_assemble_colored(int elem_type, int op, int q, const double* w, const double* pts, const double* params,
                          const double* vertices, const u64* conn,
                          u64 num_colors, const u64* color_offsets, const u64* color_elements,
                          const u64* row_offsets, const u64* col_indices, double* values, int nthreads, int64_t* bad_element) {
    oref_problem p;
    int st = init_problem(&p, elem_type, op, q, w, pts, params);
    volatile int err = 0;
    u64 c;
    if (st) return st;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = oref_max_threads();
#else
    nthreads = 1;
#endif
    for (c = 0; c < num_colors && !err; ++c) {
        int64_t b = (int64_t)color_offsets[c], e = (int64_t)color_offsets[c + 1];
#pragma omp parallel num_threads(nthreads)
        {
            double X[MAXN * MAXD];
            double* K = (double*)malloc(sizeof(double) * (size_t)(p.s * p.n) * (p.s * p.n));
            int64_t k;
#pragma omp for schedule(dynamic, 16)
            for (k = b; k < e; ++k) {
                u64 el = color_elements[k];
                const u64* nodes = conn + el * (u64)p.n;
                int s2;
                if (err) continue;
                gather_coords(&p, vertices, nodes, X);
                s2 = element_matrix(&p, X, K);
                if (!s2) s2 = scatter_element(&p, row_offsets, col_indices, values, nodes, K);
                if (s2) {
#pragma omp critical
                    { if (!err) { err = s2; if (bad_element) *bad_element = (int64_t)el; } }
                }
            }
            free(K);
        }
    }
    return err;
}

/* ---------------- pattern (global.rs:65-120), via node->element adjacency + per-node sort-unique.
 * Output identical to the hash-set formulation: per node the sorted set of coupled nodes,
 * sdim identical rows per node, columns sdim*node_j + j. Ragged connectivity (NestedVec layout):
 * elem_offsets[E+1], elem_nodes[]. */
static int cmp_u64(const void* a, const void* b) { u64 x = *(const u64*)a, y = *(const u64*)b; return (x > y) - (x < y); }

int oref_pattern(int sdim, u64 num_nodes, u64 num_elements, const u64* elem_offsets, const u64* elem_nodes_,
                 u64* row_offsets /* sdim*num_nodes+1 */, u64** col_indices_out, u64* nnz_out) {
    u64* cnt = (u64*)calloc(num_nodes + 1, sizeof(u64));
    u64* adj;
    u64* cursor;
    u64 e, k, node, total, nnz = 0, cap = 0;
    u64* buf = NULL;
    u64* blk_cnt = (u64*)calloc(num_nodes, sizeof(u64));
    u64** blk = (u64**)calloc(num_nodes ? num_nodes : 1, sizeof(u64*));
    u64* cols;
    u64 pos;
    int r, j;
    for (e = 0; e < num_elements; ++e)
        for (k = elem_offsets[e]; k < elem_offsets[e + 1]; ++k) {
            if (elem_nodes_[k] >= num_nodes) { free(cnt); free(blk_cnt); free(blk); return OREF_ERR_ARG; }
            cnt[elem_nodes_[k] + 1]++;
        }
    for (node = 0; node < num_nodes; ++node) cnt[node + 1] += cnt[node];
    total = cnt[num_nodes];
    adj = (u64*)malloc(sizeof(u64) * (total ? total : 1));
    cursor = (u64*)malloc(sizeof(u64) * (num_nodes ? num_nodes : 1));
    memcpy(cursor, cnt, sizeof(u64) * num_nodes);
    for (e = 0; e < num_elements; ++e)
        for (k = elem_offsets[e]; k < elem_offsets[e + 1]; ++k) adj[cursor[elem_nodes_[k]]++] = e;
    for (node = 0; node < num_nodes; ++node) {
        u64 m = 0, a, u;
        for (a = cnt[node]; a < cnt[node + 1]; ++a) m += elem_offsets[adj[a] + 1] - elem_offsets[adj[a]];
        if (m > cap) { cap = 2 * m; buf = (u64*)realloc(buf, sizeof(u64) * cap); }
        m = 0;
        for (a = cnt[node]; a < cnt[node + 1]; ++a)
            for (k = elem_offsets[adj[a]]; k < elem_offsets[adj[a] + 1]; ++k) buf[m++] = elem_nodes_[k];
        if (m) qsort(buf, m, sizeof(u64), cmp_u64);
        u = 0;
        for (a = 0; a < m; ++a) if (a == 0 || buf[a] != buf[a - 1]) buf[u++] = buf[a];
        blk_cnt[node] = u;
        blk[node] = (u64*)malloc(sizeof(u64) * (u ? u : 1));
        memcpy(blk[node], buf, sizeof(u64) * u);
        nnz += (u64)sdim * (u64)sdim * u;
    }
    cols = (u64*)malloc(sizeof(u64) * (nnz ? nnz : 1));
    pos = 0;
    row_offsets[0] = 0;
    for (node = 0; node < num_nodes; ++node) {
        for (r = 0; r < sdim; ++r) {
            for (k = 0; k < blk_cnt[node]; ++k)
                for (j = 0; j < sdim; ++j) cols[pos++] = (u64)sdim * blk[node][k] + (u64)j;
            row_offsets[(u64)sdim * node + (u64)r + 1] = pos;
        }
        free(blk[node]);
    }
    free(cnt); free(adj); free(cursor); free(buf); free(blk_cnt); free(blk);
    *col_indices_out = cols;
    *nnz_out = nnz;
    return OREF_OK;
}
void oref_free(void* p) { free(p); }

/* ---------------- greedy colouring (fenris-paradis/src/coloring.rs:6-70).
 * out: color_offsets (malloc, num_colors+1), color_elements (caller, E). */
int oref_color_greedy(u64 num_elements, const u64* elem_offsets, const u64* elem_nodes_, u64 num_nodes,
                      u64** color_offsets_out, u64* color_elements, u64* num_colors_out) {
    int32_t* last = (int32_t*)malloc(sizeof(int32_t) * (num_nodes ? num_nodes : 1));
    u64* cur = (u64*)malloc(sizeof(u64) * (num_elements ? num_elements : 1));
    u64* post = (u64*)malloc(sizeof(u64) * (num_elements ? num_elements : 1));
    u64 ncur = num_elements, npost, i, k, out = 0, ncol = 0, capc = 64;
    u64* offs = (u64*)malloc(sizeof(u64) * (capc + 1));
    int32_t c = 0;
    for (i = 0; i < num_nodes; ++i) last[i] = -1;
    for (i = 0; i < num_elements; ++i) cur[i] = i;
    offs[0] = 0;
    while (ncur) {
        u64* t;
        npost = 0;
        for (i = 0; i < ncur; ++i) {
            u64 e = cur[i];
            int blocked = 0;
            for (k = elem_offsets[e]; k < elem_offsets[e + 1]; ++k)
                if (last[elem_nodes_[k]] == c) { blocked = 1; break; }
            if (blocked) post[npost++] = e;
            else {
                for (k = elem_offsets[e]; k < elem_offsets[e + 1]; ++k) last[elem_nodes_[k]] = c;
                color_elements[out++] = e;
            }
        }
        if (ncol + 1 > capc) { capc *= 2; offs = (u64*)realloc(offs, sizeof(u64) * (capc + 1)); }
        offs[++ncol] = out;
        t = cur; cur = post; post = t;
        ncur = npost;
        ++c;
    }
    free(last); free(cur); free(post);
    *color_offsets_out = offs;
    *num_colors_out = ncol;
    return OREF_OK;
}

/* ---------------- mesh generators (src/mesh/procedural.rs) ---------------- */
/* :216-277. vertices (n+1)^3 x 3, conn n^3 x 8 */
void oref_gen_hex_mesh(u64 cx, u64 cy, u64 cz, double cell_size, double* vertices, u64* conn) {
    u64 vx = cx + 1, vy = cy + 1, vz = cz + 1, i, j, k, p = 0;
    for (k = 0; k < vz; ++k) for (j = 0; j < vy; ++j) for (i = 0; i < vx; ++i) {
        vertices[p++] = (double)i * cell_size; vertices[p++] = (double)j * cell_size; vertices[p++] = (double)k * cell_size;
    }
#define HIDX(i, j, k) ((vx * vy) * (k) + vx * (j) + (i))
    p = 0;
    for (k = 0; k < cz; ++k) for (j = 0; j < cy; ++j) for (i = 0; i < cx; ++i) {
        conn[p++] = HIDX(i, j, k); conn[p++] = HIDX(i + 1, j, k); conn[p++] = HIDX(i + 1, j + 1, k); conn[p++] = HIDX(i, j + 1, k);
        conn[p++] = HIDX(i, j, k + 1); conn[p++] = HIDX(i + 1, j, k + 1); conn[p++] = HIDX(i + 1, j + 1, k + 1); conn[p++] = HIDX(i, j + 1, k + 1);
    }
#undef HIDX
}

/* :46-93 with top_left = (0, 1) */
void oref_gen_quad_mesh(u64 cells, double cell_size, double* vertices, u64* conn) {
    u64 nx = cells, ny = cells, i, j, p = 0;
    for (j = 0; j <= ny; ++j) for (i = 0; i <= nx; ++i) {
        vertices[p++] = 0.0 + (double)i * cell_size; vertices[p++] = 1.0 + (-(double)j) * cell_size;
    }
#define QIDX(i, j) ((nx + 1) * (j) + (i))
    p = 0;
    for (j = 0; j < ny; ++j) for (i = 0; i < nx; ++i) {
        conn[p++] = QIDX(i, j + 1); conn[p++] = QIDX(i + 1, j + 1); conn[p++] = QIDX(i + 1, j); conn[p++] = QIDX(i, j);
    }
#undef QIDX
}

/* :286-403. vertices ((c+1)^3 + c^3) x 3, conn 12 c^3 x 4 (for a box of cx*cy*cz cells) */
static const int PFD[3][4][3] = {
    {{1, 0, 1}, {1, 1, 1}, {1, 1, 0}, {1, 0, 0}},
    {{0, 1, 0}, {1, 1, 0}, {1, 1, 1}, {0, 1, 1}},
    {{0, 1, 1}, {1, 1, 1}, {1, 0, 1}, {0, 0, 1}}};

u64 oref_gen_tet_mesh(u64 cx, u64 cy, u64 cz, double cell_size, double* vertices, u64* conn) {
    u64 vx = cx + 1, vy = cy + 1, vz = cz + 1, i, j, k, p = 0, off, t = 0;
    u64 nc[3];
    int axis, m, pos;
    nc[0] = cx; nc[1] = cy; nc[2] = cz;
    for (k = 0; k < vz; ++k) for (j = 0; j < vy; ++j) for (i = 0; i < vx; ++i) {
        vertices[p++] = cell_size * (double)i; vertices[p++] = cell_size * (double)j; vertices[p++] = cell_size * (double)k;
    }
    off = vx * vy * vz;
    for (k = 0; k < cz; ++k) for (j = 0; j < cy; ++j) for (i = 0; i < cx; ++i) {
        vertices[p++] = cell_size * (0.5 + (double)i); vertices[p++] = cell_size * (0.5 + (double)j); vertices[p++] = cell_size * (0.5 + (double)k);
    }
#define VIDX(a, b, c) ((vx * vy) * (u64)(c) + vx * (u64)(b) + (u64)(a))
#define CIDX(a, b, c) ((cx * cy) * (u64)(c) + cx * (u64)(b) + (u64)(a) + off)
    for (k = 0; k < cz; ++k) for (j = 0; j < cy; ++j) for (i = 0; i < cx; ++i) {
        u64 cell[3];
        cell[0] = i; cell[1] = j; cell[2] = k;
        for (axis = 0; axis < 3; ++axis) {
            if (cell[axis] + 1 < nc[axis]) {
                u64 face[4], c1, c2, nb[3];
                for (m = 0; m < 4; ++m) face[m] = VIDX(i + PFD[axis][m][0], j + PFD[axis][m][1], k + PFD[axis][m][2]);
                c1 = CIDX(i, j, k);
                nb[0] = i; nb[1] = j; nb[2] = k; nb[axis] += 1;
                c2 = CIDX(nb[0], nb[1], nb[2]);
                for (m = 0; m < 4; ++m) {
                    u64 v1 = face[m], v2 = face[(m + 1) % 4];
                    conn[t++] = c1; conn[t++] = c2; conn[t++] = v2; conn[t++] = v1;
                }
            }
            for (pos = 0; pos < 2; ++pos) {
                if ((pos == 0 && cell[axis] == 0) || (pos == 1 && cell[axis] + 1 == nc[axis])) {
                    int64_t fv[4][3];
                    u64 a, b, c, d, center;
                    for (m = 0; m < 4; ++m) {
                        fv[m][0] = (int64_t)i + PFD[axis][m][0]; fv[m][1] = (int64_t)j + PFD[axis][m][1]; fv[m][2] = (int64_t)k + PFD[axis][m][2];
                    }
                    if (pos == 0) {
                        int64_t tmp[3];
                        for (m = 0; m < 2; ++m) { memcpy(tmp, fv[m], sizeof tmp); memcpy(fv[m], fv[3 - m], sizeof tmp); memcpy(fv[3 - m], tmp, sizeof tmp); }
                        for (m = 0; m < 4; ++m) fv[m][axis] -= 1;
                    }
                    a = VIDX(fv[0][0], fv[0][1], fv[0][2]); b = VIDX(fv[1][0], fv[1][1], fv[1][2]);
                    c = VIDX(fv[2][0], fv[2][1], fv[2][2]); d = VIDX(fv[3][0], fv[3][1], fv[3][2]);
                    center = CIDX(i, j, k);
                    if ((i + j + k) % 2 == 0) {
                        conn[t++] = a; conn[t++] = b; conn[t++] = c; conn[t++] = center;
                        conn[t++] = a; conn[t++] = c; conn[t++] = d; conn[t++] = center;
                    } else {
                        conn[t++] = a; conn[t++] = b; conn[t++] = d; conn[t++] = center;
                        conn[t++] = b; conn[t++] = c; conn[t++] = d; conn[t++] = center;
                    }
                }
            }
        }
    }
#undef VIDX
#undef CIDX
    return t / 4;
}
