/* The assembly path from plain C, through the C ABI only (include/fenris_b200.h) - what a cgo / Rust -sys / JNI binding would call.
 *
 *   gcc -std=c99 -O2 -I include examples/assemble_csr.c -L fenris_b200/lib -lfenris_b200 -Wl,-rpath,$PWD/fenris_b200/lib -lm -o build/assemble_csr
 *   build/assemble_csr [cells per side = 16]
 *
 * Reference call sequence this replaces (Rust): create_unit_box_uniform_hex_mesh_3d(n) (src/mesh/procedural.rs:216-277),
 * CsrAssembler::assemble_pattern (src/assembly/global.rs:65-120), ElementEllipticAssembler over MaterialEllipticOperator<LinearElasticMaterial>
 * with the canonical Hex8 rule, CsrAssembler::assemble (global.rs:124-182).
 * The program checks what holds for every mesh without an oracle: K is symmetric (entry by entry, through the downloaded pattern) and
 * rigid translations lie in its null space (row sums per component vanish).  Exit code 0 = checks passed, 2 = a check failed,
 * 3 = the library reported an error (without a GPU: fb200_create fails with FB200_ERR_CUDA - there is no CPU fallback).
 * tests/test_host.py compiles and links it on every CPU run. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fenris_b200.h"

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        fb200_status st_ = (call);                                                                    \
        if (st_ != FB200_OK) {                                                                        \
            char msg[256] = "";                                                                       \
            int64_t elem = -1;                                                                        \
            if (ctx) fb200_last_error(ctx, msg, sizeof msg, &elem);                                   \
            fprintf(stderr, "%s -> %s %s\n", #call, fb200_status_string(st_), msg);                   \
            return 3;                                                                                 \
        }                                                                                             \
    } while (0)

int main(int argc, char** argv) {
    fb200_ctx* ctx = NULL;
    const uint64_t n = argc > 1 ? strtoull(argv[1], NULL, 10) : 16;
    if (fb200_abi_version() != FB200_VERSION) {
        fprintf(stderr, "header / library version mismatch\n");
        return 3;
    }
    /* mesh on the host (the caller's Mesh<f64, U3, Hex8Connectivity>) */
    uint64_t nv = 0, ne = 0;
    CHECK(fb200_gen_hex_mesh(n, n, n, 1.0 / (double)n, &nv, &ne, NULL, NULL));
    double* vertices = malloc(sizeof(double) * 3 * nv);
    uint64_t* conn = malloc(sizeof(uint64_t) * 8 * ne);
    if (!vertices || !conn) return 3;
    CHECK(fb200_gen_hex_mesh(n, n, n, 1.0 / (double)n, &nv, &ne, vertices, conn));
    /* canonical rule + uniform Lame data (UniformQuadratureTable::with_uniform_data) */
    int32_t nq = 0;
    CHECK(fb200_canonical_quadrature(FB200_HEX8, &nq, NULL, NULL));
    double* w = malloc(sizeof(double) * nq);
    double* pts = malloc(sizeof(double) * 3 * nq);
    double* data = malloc(sizeof(double) * 2 * nq);
    if (!w || !pts || !data) return 3;
    CHECK(fb200_canonical_quadrature(FB200_HEX8, &nq, w, pts));
    double mu, lambda;
    fb200_lame_from_young_poisson(1e6, 0.2, &mu, &lambda);
    for (int32_t q = 0; q < nq; ++q) {
        data[2 * q] = mu;
        data[2 * q + 1] = lambda;
    }
    const fb200_quadrature rule = {nq, 3, w, pts, data};
    const fb200_operator op = {FB200_LINEAR_ELASTIC};

    CHECK(fb200_create(0, &ctx));
    CHECK(fb200_space_upload(ctx, FB200_HEX8, nv, vertices, ne, conn));
    uint64_t rows = 0, nnz = 0;
    CHECK(fb200_assemble_pattern(ctx, 3, &rows, &nnz));
    uint64_t* offsets = malloc(sizeof(uint64_t) * (rows + 1));
    uint64_t* cols = malloc(sizeof(uint64_t) * nnz);
    double* values = malloc(sizeof(double) * nnz);
    if (!offsets || !cols || !values) return 3;
    CHECK(fb200_pattern_download(ctx, offsets, cols));
    CHECK(fb200_timer_begin(ctx));
    CHECK(fb200_assemble_into_csr_device(ctx, &op, &rule, NULL, FB200_SCATTER_ATOMIC, 0));
    float ms = 0.f;
    CHECK(fb200_timer_end(ctx, &ms));
    CHECK(fb200_synchronize(ctx)); /* deferred kernel errors (singular Jacobian ...) */
    CHECK(fb200_values_download(ctx, values));

    /* checks: null space of the translations, symmetry */
    double vmax = 0.0, worst_sum = 0.0, worst_asym = 0.0;
    for (uint64_t k = 0; k < nnz; ++k) vmax = fmax(vmax, fabs(values[k]));
    for (uint64_t r = 0; r < rows; ++r) {
        double sum[3] = {0.0, 0.0, 0.0};
        for (uint64_t k = offsets[r]; k < offsets[r + 1]; ++k) {
            sum[cols[k] % 3] += values[k];
            /* the transposed entry: binary search of r in row cols[k] */
            const uint64_t c = cols[k];
            uint64_t lo = offsets[c], hi = offsets[c + 1];
            while (lo < hi) {
                const uint64_t mid = lo + (hi - lo) / 2;
                if (cols[mid] < r) lo = mid + 1;
                else hi = mid;
            }
            if (lo == offsets[c + 1] || cols[lo] != r) {
                fprintf(stderr, "pattern not symmetric at (%llu, %llu)\n", (unsigned long long)r, (unsigned long long)c);
                return 2;
            }
            worst_asym = fmax(worst_asym, fabs(values[k] - values[lo]));
        }
        for (int j = 0; j < 3; ++j) worst_sum = fmax(worst_sum, fabs(sum[j]));
    }
    printf("Hex8 elasticity %llu^3 cells: %llu elements, %llu rows, nnz %llu, assembly %.3f ms (%.3g elements/s), %llu kernel launches\n",
           (unsigned long long)n, (unsigned long long)ne, (unsigned long long)rows, (unsigned long long)nnz, ms, (double)ne / (1e-3 * ms),
           (unsigned long long)fb200_launch_count(ctx));
    printf("max |K_ij| %.6e, max |K_ij - K_ji| %.3e, max |row sum per component| %.3e\n", vmax, worst_asym, worst_sum);
    const int ok = worst_asym <= 1e-12 * vmax && worst_sum <= 1e-10 * vmax;
    fb200_destroy(ctx);
    free(values), free(cols), free(offsets), free(data), free(pts), free(w), free(conn), free(vertices);
    return ok ? 0 : 2;
}
