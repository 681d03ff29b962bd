// Measured device ceilings for the roofline statements of bench.py ("FP64-pipe utilisation vs a MEASURED
// FP64 FMA peak -- MEASURED_PEAKS.json has no fp64 entry").  Nothing on the product path calls these kernels.
//
//   k_fp64_fma : every thread runs 8 independent dependent-FMA chains (enough ILP to cover the DFMA latency at the
//                residency used), 2 flops per FMA; grid = SMs x resident blocks, so the figure is the whole-chip
//                double-precision FMA issue rate at the clocks the run actually had.
//   k_l2_read  : grid-stride 16-byte loads over a buffer that fits the 126 MB L2 (default 32 MiB), read repeatedly:
//                the L2 -> SM bandwidth the mesh gathers of the path are served from.
#include "common.cuh"
#include "kernels.h"

namespace css {

namespace {

__global__ void __launch_bounds__(256) k_fp64_fma(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0 - 1e-9, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c), a1 = fma(a1, m, c), a2 = fma(a2, m, c), a3 = fma(a3, m, c);
            a4 = fma(a4, m, c), a5 = fma(a5, m, c), a6 = fma(a6, m, c), a7 = fma(a7, m, c);
        }
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.6789) out[blockIdx.x] = s; // keeps the chains alive; practically never true
}

__global__ void __launch_bounds__(256) k_l2_read(const int4* buf, size_t n16, int passes, int* out)
{
    int acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            int4 v = __ldcg(buf + i);
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x5a5a5a5a) out[0] = acc;
}

} // namespace

// returns TFLOP/s (what = 0) or GB/s (what = 1), best of `reps` event-timed launches after one warm-up; < 0 on error
double runMicrobench(cudaStream_t st, int numSMs, int what, int reps)
{
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1;
    double best = -1;
    if (what == 0) {
        const int blocks = numSMs * 8, threads = 256, iters = 4096;
        double* out = nullptr;
        if (cudaMalloc(&out, sizeof(double) * blocks) != cudaSuccess) return -1;
        for (int r = 0; r <= reps; ++r) {
            cudaEventRecord(e0, st);
            k_fp64_fma<<<blocks, threads, 0, st>>>(out, iters, 0.5);
            cudaEventRecord(e1, st);
            if (cudaEventSynchronize(e1) != cudaSuccess) break;
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
            if (r > 0 && ms > 0) best = fmax(best, flops / (ms * 1e-3) / 1e12);
        }
        cudaFree(out);
    } else {
        const size_t bytes = 32ull << 20, n16 = bytes / 16;
        const int passes = 16;
        int4* buf = nullptr;
        int* out = nullptr;
        if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&out, 4) != cudaSuccess) return -1;
        cudaMemsetAsync(buf, 1, bytes, st);
        for (int r = 0; r <= reps; ++r) {
            cudaEventRecord(e0, st);
            k_l2_read<<<numSMs * 8, 256, 0, st>>>(buf, n16, passes, out);
            cudaEventRecord(e1, st);
            if (cudaEventSynchronize(e1) != cudaSuccess) break;
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms > 0) best = fmax(best, (double)bytes * passes / (ms * 1e-3) / 1e9);
        }
        cudaFree(buf), cudaFree(out);
    }
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    return cudaGetLastError() == cudaSuccess ? best : -1;
}

} // namespace css
