// Scatter-path microbenchmark for B200 (sm_100a): what bounds fp64 reductions / stores into a large CSR value array?
// Every warp issues warp-wide red.global.add.f64 (or st.global / TMA bulk reductions) with a chosen lane -> address pattern
// into a window that is either L2 resident (16 MB) or streaming (4 GB: every line is a DRAM fill).  Reported per pattern:
// warp instructions/s, lanes/s, 32 B sector-ops/s and 128 B line-packets/s, so that the cost model of the assembly scatter
//  rests on measurements.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/microbench_red scripts/microbench_red.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

enum Pattern { CONTIG24 = 0, RUNS4X6, RUNS8X3, CONTIG32, SCATTER32, RUNS4X6_J0, NPATTERN };
static const char* kPatName[] = {"contig24", "runs4x6", "runs8x3", "contig32", "scatter32", "runs4x6_j0"};
// (active lanes, sectors, lines) per instruction for the report
static const int kLanes[] = {24, 24, 24, 32, 32, 8};
static const double kSectors[] = {6.75, 9.0, 8.0, 8.0, 32.0, 7.0};
static const double kLines[] = {2.0, 4.0, 8.0, 2.0, 32.0, 4.0};

__device__ __forceinline__ int lane_offset(int pattern, int lane, bool& active) {
    active = true;
    switch (pattern) {
        case CONTIG24: active = lane < 24; return lane + 1;
        case RUNS4X6: active = lane < 24; return (lane / 6) * 32 + lane % 6 + 1;
        case RUNS8X3: active = lane < 24; return (lane / 3) * 16 + lane % 3 + 1;
        case CONTIG32: return lane;
        case SCATTER32: return lane * 16;
        case RUNS4X6_J0: active = lane < 24 && (lane % 3) == 0; return (lane / 6) * 32 + lane % 6 + 1;
    }
    return lane;
}

template <int MODE>  // 0 red, 1 st, 2 ld + st
__global__ void __launch_bounds__(128) scatter_kernel(double* base, uint64_t window_mask, int iters, int pattern, uint64_t region_stride) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    bool active;
    const int off = lane_offset(pattern, lane, active);
    for (int it = 0; it < iters; ++it) {
        const uint64_t r0 = ((uint64_t)it * nwarps + warp) * 8;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            // regions are 4 KB apart (512 doubles): the 8 instructions of one iteration behave like 8 different CSR rows
            double* dst = base + (((r0 + u) * region_stride) & window_mask) + off;
            if (active) {
                if (MODE == 0) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst), "d"(1.0) : "memory");
                else if (MODE == 1) *reinterpret_cast<volatile double*>(dst) = 1.0;
                else *dst += 1.0;
            }
        }
    }
}

// TMA bulk reduction: lane 0 of every warp adds `bytes` contiguous bytes from shared memory to global memory
__global__ void __launch_bounds__(128) bulk_kernel(double* base, uint64_t window_mask, int iters, int bytes, uint64_t region_stride) {
    extern __shared__ __align__(128) double sm[];
    for (int i = threadIdx.x; i < 4 * 512; i += blockDim.x) sm[i] = 1.0;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;");
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(sm + wl * 512);
    if (lane != 0) return;
    for (int it = 0; it < iters; ++it) {
        const uint64_t r0 = ((uint64_t)it * nwarps + warp) * 8;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double* dst = base + (((r0 + u) * region_stride) & window_mask);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;");
        asm volatile("cp.async.bulk.wait_group.read 2;");
    }
    asm volatile("cp.async.bulk.wait_group 0;");
}

static double time_ms(cudaEvent_t e0, cudaEvent_t e1) {
    float ms = 0;
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const uint64_t big = 1ull << 29;  // 4 GB of doubles
    double* buf;
    if (cudaMalloc(&buf, (big + 4096) * sizeof(double)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, (big + 4096) * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 400;
    for (int warps_per_sm : {8, 16, 32}) {
        const int blocks = sms * warps_per_sm / 4;
        const double ninstr = (double)blocks * 4 * iters * 8;
        for (int resident = 1; resident >= 0; --resident) {
            const uint64_t mask = (resident ? (1ull << 21) : big) - 1;  // 16 MB window or 4 GB
            for (int mode = 0; mode < 3; ++mode) {
                for (int pat = 0; pat < NPATTERN; ++pat) {
                    auto launch = [&](int n) {
                        if (mode == 0) scatter_kernel<0><<<blocks, 128>>>(buf, mask, n, pat, 512);
                        else if (mode == 1) scatter_kernel<1><<<blocks, 128>>>(buf, mask, n, pat, 512);
                        else scatter_kernel<2><<<blocks, 128>>>(buf, mask, n, pat, 512);
                    };
                    launch(20);
                    cudaDeviceSynchronize();
                    cudaEventRecord(e0);
                    launch(iters);
                    cudaEventRecord(e1);
                    const double ms = time_ms(e0, e1);
                    const double ips = ninstr / (ms * 1e-3);
                    printf("%-6s %-9s w/SM=%2d %-10s %8.3f ms  instr/s %.3e (%.2f cyc/instr/SM)  lanes/s %.3e  sectors/s %.3e  lines/s %.3e  payload %.0f GB/s\n",
                           mode == 0 ? "red" : (mode == 1 ? "st" : "ld+st"), resident ? "L2-16MB" : "DRAM-4GB", warps_per_sm, kPatName[pat], ms, ips,
                           1.965e9 * sms / ips, ips * kLanes[pat], ips * kSectors[pat], ips * kLines[pat], ips * kLanes[pat] * 8 / 1e9);
                }
            }
        }
    }
    // TMA bulk reductions
    for (int warps_per_sm : {4, 16}) {
        const int blocks = sms * warps_per_sm / 4;
        const double nops = (double)blocks * 4 * iters * 8;
        for (int resident = 1; resident >= 0; --resident) {
            const uint64_t mask = (resident ? (1ull << 21) : big) - 1;
            for (int bytes : {32, 48, 96, 192, 656, 2048}) {
                bulk_kernel<<<blocks, 128, 4 * 512 * sizeof(double)>>>(buf, mask, 20, bytes, 512);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                bulk_kernel<<<blocks, 128, 4 * 512 * sizeof(double)>>>(buf, mask, iters, bytes, 512);
                cudaEventRecord(e1);
                const double ms = time_ms(e0, e1);
                const double ops = nops / (ms * 1e-3);
                printf("bulkred %-9s w/SM=%2d bytes=%4d %8.3f ms  ops/s %.3e (%.2f cyc/op/SM)  sectors/s %.3e  payload %.0f GB/s\n", resident ? "L2-16MB" : "DRAM-4GB",
                       warps_per_sm, bytes, ms, ops, 1.965e9 * sms / ops, ops * bytes / 32.0, ops * bytes / 1e9);
            }
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
