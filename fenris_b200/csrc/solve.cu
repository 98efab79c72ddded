// the immediate consumer of the assembled matrix - y = A x and a (Jacobi-)preconditioned conjugate gradient
// on the DEVICE-RESIDENT CSR, so that a multi-GB matrix never has to cross PCIe to be used.
//
// Reference semantics: ConjugateGradient::solve_with_guess (fenris-sparse/src/cg.rs:364-480) with RelativeResidualCriterion
// (cg.rs:85-124: ||r|| <= tol ||b||, the recursively updated residual), errors IndefiniteOperator (p.Ap <= 0),
// IndefinitePreconditioner (z.r <= 0), MaxIterationsReached; preconditioner = identity or the inverse diagonal (the callers of the
// reference pass an arbitrary LinearOperator; Jacobi is what its tests use, e.g. poisson_mms_common.rs:142-164 uses plain CG).
//
// The matrix is the node-block CSR of pattern.cu: node I owns s consecutive rows that share one column pattern (global.rs:86-109).
// y = A x: one warp per node, lanes over the coupled nodes, s partial sums per lane, xor-shuffle reduction - every value is read once,
// fully coalesced (the s x cnt row panel is contiguous); dot products are two-stage and deterministic.
#include <cmath>

#include "fb200_internal.h"

namespace fb200 {

template <int S>
__global__ void __launch_bounds__(256) spmv_kernel(uint64_t num_nodes, const int64_t* __restrict__ blk_off, const int32_t* __restrict__ blk_cols,
                                                  const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t I = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; I < num_nodes; I += nwarps) {
        const int64_t b = blk_off[I], cnt = blk_off[I + 1] - b;
        const double* v = values + (int64_t)(S * S) * b;
        double acc[S];
#pragma unroll
        for (int i = 0; i < S; ++i) acc[i] = 0.0;
        // lanes run over the scalar columns c = S k + j of the row panel: consecutive lanes read consecutive values
        for (int64_t c = lane; c < S * cnt; c += 32) {
            const int64_t k = c / S;
            const double xv = x[(int64_t)blk_cols[b + k] * S + (c - k * S)];
#pragma unroll
            for (int i = 0; i < S; ++i) acc[i] = fma(v[(int64_t)i * S * cnt + c], xv, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < S; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        }
        if (lane < S) y[I * S + lane] = acc[lane];
    }
}

// diag[r] of the node-block CSR (0 where the pattern has no diagonal entry)
__global__ void diagonal_kernel(uint64_t num_nodes, const int64_t* __restrict__ blk_off, const int32_t* __restrict__ blk_cols, int s,
                                const double* __restrict__ values, double* __restrict__ diag) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t I = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; I < num_nodes; I += stride) {
        const int64_t b = blk_off[I], e = blk_off[I + 1];
        int64_t lo = b, hi = e;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (blk_cols[mid] < (int32_t)I) lo = mid + 1; else hi = mid;
        }
        const bool has = lo < e && blk_cols[lo] == (int32_t)I;
        for (int i = 0; i < s; ++i) diag[I * s + i] = has ? values[(int64_t)(s * s) * b + (int64_t)i * s * (e - b) + s * (lo - b) + i] : 0.0;
    }
}

constexpr int kDotBlocks = 1024;
// partial[blockIdx] = sum over a grid-stride slice of a_i * b_i ; a second launch with one block adds the partials (fixed order)
__global__ void __launch_bounds__(256) dot_partial_kernel(const double* __restrict__ a, const double* __restrict__ b, uint64_t n, double* partial) {
    __shared__ double sh[256];
    double s = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) s = fma(a[i], b[i], s);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) dot_final_kernel(const double* __restrict__ partial, int count, double* out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += 256) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

// elementwise helpers of the CG loop (cg.rs:381-474)
__global__ void residual_kernel(double* r, const double* __restrict__ b, uint64_t n) {  // r <- b - r
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) r[i] = b[i] - r[i];
}
__global__ void precondition_kernel(double* z, const double* __restrict__ r, const double* __restrict__ diag, int jacobi, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        z[i] = jacobi ? (diag[i] != 0.0 ? r[i] / diag[i] : r[i]) : r[i];
}
__global__ void update_xr_kernel(double* x, double* r, const double* __restrict__ p, const double* __restrict__ Ap, double alpha, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        x[i] += alpha * p[i];
        r[i] -= alpha * Ap[i];
    }
}
__global__ void update_p_kernel(double* p, const double* __restrict__ z, double beta, uint64_t n) {  // p <- beta p + z
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double v = p[i];
        v *= beta;
        v += z[i];
        p[i] = v;
    }
}

static fb200_status spmv_device(fb200_ctx* ctx, const double* d_x, double* d_y) {
    if (ctx->N == 0) return FB200_OK;
    const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(ctx->N * 32, 256), (uint64_t)ctx->sm_count * 16));
    switch (ctx->sdim) {
        case 1: spmv_kernel<1><<<blocks, 256, 0, ctx->stream>>>(ctx->N, ctx->d_blk_off, ctx->d_blk_cols, ctx->d_values, d_x, d_y); break;
        case 2: spmv_kernel<2><<<blocks, 256, 0, ctx->stream>>>(ctx->N, ctx->d_blk_off, ctx->d_blk_cols, ctx->d_values, d_x, d_y); break;
        default: spmv_kernel<3><<<blocks, 256, 0, ctx->stream>>>(ctx->N, ctx->d_blk_off, ctx->d_blk_cols, ctx->d_values, d_x, d_y); break;
    }
    return check_launch(ctx, "spmv_kernel");
}

struct CgWork {
    double *x = nullptr, *b = nullptr, *r = nullptr, *z = nullptr, *p = nullptr, *Ap = nullptr, *diag = nullptr, *partial = nullptr, *scalar = nullptr;
    ~CgWork() {
        for (double* q : {x, b, r, z, p, Ap, diag, partial, scalar})
            if (q) cudaFree(q);
    }
};

static fb200_status dot_device(fb200_ctx* ctx, CgWork& w, const double* a, const double* b, uint64_t n, double* host_out) {
    dot_partial_kernel<<<kDotBlocks, 256, 0, ctx->stream>>>(a, b, n, w.partial);
    FB200_TRY(check_launch(ctx, "dot_partial_kernel"));
    dot_final_kernel<<<1, 256, 0, ctx->stream>>>(w.partial, kDotBlocks, w.scalar);
    FB200_TRY(check_launch(ctx, "dot_final_kernel"));
    FB200_CUDA(ctx, cudaMemcpyAsync(host_out, w.scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FB200_OK;
}

}  // namespace fb200

using namespace fb200;

extern "C" {

fb200_status fb200_spmv(fb200_ctx* ctx, const double* x, double* y) {
    if (!ctx) return FB200_ERR_STATE;
    if (!ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no matrix: assemble first");
    if (!x || !y) return fail(ctx, FB200_ERR_SHAPE, "null vector");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = ctx->nrows;
    CgWork w;
    FB200_TRY(dev_alloc(ctx, &w.x, n));
    FB200_TRY(dev_alloc(ctx, &w.Ap, n));
    if (n) FB200_CUDA(ctx, cudaMemcpyAsync(w.x, x, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FB200_TRY(spmv_device(ctx, w.x, w.Ap));
    if (n) FB200_CUDA(ctx, cudaMemcpyAsync(y, w.Ap, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FB200_OK;
}

fb200_status fb200_cg_solve(fb200_ctx* ctx, const double* b, double* x, double rel_tol, uint64_t max_iter, int32_t jacobi, uint64_t* iterations,
                            double* rel_residual) {
    if (!ctx) return FB200_ERR_STATE;
    if (!ctx->has_pattern) return fail(ctx, FB200_ERR_STATE, "no matrix: assemble first");
    if (!b || !x) return fail(ctx, FB200_ERR_SHAPE, "null vector");
    FB200_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = ctx->nrows;
    if (iterations) *iterations = 0;
    if (rel_residual) *rel_residual = 0.0;
    if (n == 0) return FB200_OK;
    CgWork w;
    FB200_TRY(dev_alloc(ctx, &w.x, n));
    FB200_TRY(dev_alloc(ctx, &w.b, n));
    FB200_TRY(dev_alloc(ctx, &w.r, n));
    FB200_TRY(dev_alloc(ctx, &w.z, n));
    FB200_TRY(dev_alloc(ctx, &w.p, n));
    FB200_TRY(dev_alloc(ctx, &w.Ap, n));
    FB200_TRY(dev_alloc(ctx, &w.diag, n));
    FB200_TRY(dev_alloc(ctx, &w.partial, kDotBlocks));
    FB200_TRY(dev_alloc(ctx, &w.scalar, 1));
    const int vb = (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(n, 256), (uint64_t)ctx->sm_count * 16));
    FB200_CUDA(ctx, cudaMemcpyAsync(w.x, x, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FB200_CUDA(ctx, cudaMemcpyAsync(w.b, b, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (jacobi) {
        diagonal_kernel<<<(int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(ctx->N, 256), (uint64_t)ctx->sm_count * 16)), 256, 0, ctx->stream>>>(
            ctx->N, ctx->d_blk_off, ctx->d_blk_cols, ctx->sdim, ctx->d_values, w.diag);
        FB200_TRY(check_launch(ctx, "diagonal_kernel"));
    }
    // r = b - A x ; z = P r ; p = z   (cg.rs:381-397)
    FB200_TRY(spmv_device(ctx, w.x, w.r));
    residual_kernel<<<vb, 256, 0, ctx->stream>>>(w.r, w.b, n);
    FB200_TRY(check_launch(ctx, "residual_kernel"));
    precondition_kernel<<<vb, 256, 0, ctx->stream>>>(w.z, w.r, w.diag, jacobi, n);
    FB200_TRY(check_launch(ctx, "precondition_kernel"));
    FB200_CUDA(ctx, cudaMemcpyAsync(w.p, w.z, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    double zTr = 0.0, bb = 0.0, rr = 0.0, pAp = 0.0;
    FB200_TRY(dot_device(ctx, w, w.z, w.r, n, &zTr));
    FB200_TRY(dot_device(ctx, w, w.b, w.b, n, &bb));
    const double b_norm = std::sqrt(bb);
    fb200_status result = FB200_OK;
    uint64_t it = 0;
    if (b_norm == 0.0) {  // cg.rs:403-406
        FB200_CUDA(ctx, cudaMemsetAsync(w.x, 0, n * sizeof(double), ctx->stream));
    } else {
        for (;;) {
            FB200_TRY(dot_device(ctx, w, w.r, w.r, n, &rr));
            if (std::sqrt(rr) <= rel_tol * b_norm) break;  // RelativeResidualCriterion, cg.rs:107-124
            if (max_iter && it >= max_iter) {
                result = fail(ctx, FB200_ERR_NOT_CONVERGED, "conjugate gradient: maximum number of iterations reached");
                break;
            }
            FB200_TRY(spmv_device(ctx, w.p, w.Ap));
            FB200_TRY(dot_device(ctx, w, w.p, w.Ap, n, &pAp));
            if (!(pAp > 0.0)) {
                result = fail(ctx, FB200_ERR_INDEFINITE, "conjugate gradient: indefinite operator (p.Ap <= 0)");
                break;
            }
            if (!(zTr > 0.0)) {
                result = fail(ctx, FB200_ERR_INDEFINITE, "conjugate gradient: indefinite preconditioner (z.r <= 0)");
                break;
            }
            const double alpha = zTr / pAp;
            update_xr_kernel<<<vb, 256, 0, ctx->stream>>>(w.x, w.r, w.p, w.Ap, alpha, n);
            FB200_TRY(check_launch(ctx, "update_xr_kernel"));
            ++it;
            precondition_kernel<<<vb, 256, 0, ctx->stream>>>(w.z, w.r, w.diag, jacobi, n);
            FB200_TRY(check_launch(ctx, "precondition_kernel"));
            double zTr_next = 0.0;
            FB200_TRY(dot_device(ctx, w, w.z, w.r, n, &zTr_next));
            update_p_kernel<<<vb, 256, 0, ctx->stream>>>(w.p, w.z, zTr_next / zTr, n);
            FB200_TRY(check_launch(ctx, "update_p_kernel"));
            zTr = zTr_next;
        }
    }
    FB200_CUDA(ctx, cudaMemcpyAsync(x, w.x, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (iterations) *iterations = it;
    if (rel_residual) *rel_residual = b_norm > 0.0 ? std::sqrt(rr) / b_norm : 0.0;
    return result;
}

}  // extern "C"
