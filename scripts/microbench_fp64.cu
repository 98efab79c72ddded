// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync m8n8k4 / m16n8k4 / m16n8k8 .f64) warp-instruction
// throughput, used to decide whether the node-block contraction S = G G^T belongs on DMMA.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/microbench_fp64 scripts/microbench_fp64.cu
#include <cuda_runtime.h>

#include <cstdio>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int KIND>
__global__ void __launch_bounds__(128) bench_kernel(double* out, int iters, double seed) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = seed * (i + 1) + threadIdx.x;
    const double a = seed + 1.0, b = seed * 0.5;
    if (KIND == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
        }
    } else if (KIND == 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(acc[2 * i], acc[2 * i + 1], a, b);
        }
    } else {
        const double av[4] = {a, b, a + b, a - b};
        const double bv[2] = {b, a};
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double(&d)[4] = *reinterpret_cast<double(*)[4]>(&acc[4 * i]);
                dmma1688(d, av, bv);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
static void run(const char* name, int per_iter_instr, double flop_per_instr, int blocks_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * blocks_per_sm, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * 128);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    bench_kernel<KIND><<<blocks, 128>>>(out, 100, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench_kernel<KIND><<<blocks, 128>>>(out, iters, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)blocks * 4 * iters * per_iter_instr;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * khz * 1e3;
    printf("%-10s warps/SMSP=%d  %.3f ms  %.2f TFLOP/s  cycles per warp-instr per SMSP (at %d kHz nominal) = %.2f\n", name, blocks_per_sm, ms,
           instr * flop_per_instr / (ms * 1e-3) / 1e12, khz, cyc / (instr / (sms * 4.0)));
    cudaFree(out);
}

int main() {
    for (int bps : {1, 2, 4}) {
        run<0>("DFMA", 16, 64.0, bps);
        run<1>("DMMA884", 8, 512.0, bps);
        run<2>("DMMA1688", 4, 2048.0, bps);
    }
    return 0;
}
