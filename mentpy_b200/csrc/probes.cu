// Measurement probes used by bench.py for the second roofline ("measure the FP64 peak
// with an FMA microbenchmark"): a register-only chain of independent DFMAs per thread, and a
// STREAM-style 16-byte copy.  Not on the product path.
#include "host_util.h"

namespace {

// 8 independent accumulators per thread, `iters` rounds of 8 DFMA each: 16 flops per round-lane.
__global__ void __launch_bounds__(256) fp64_fma_probe_kernel(int64_t iters, double a, double b, double* __restrict__ out) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3;
    double x4 = x0 + 4e-3, x5 = x0 + 5e-3, x6 = x0 + 6e-3, x7 = x0 + 7e-3;
#pragma unroll 4
    for (int64_t i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b);
        x1 = fma(x1, a, b);
        x2 = fma(x2, a, b);
        x3 = fma(x3, a, b);
        x4 = fma(x4, a, b);
        x5 = fma(x5, a, b);
        x6 = fma(x6, a, b);
        x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.6789) out[0] = s;  // keeps the chain alive, never true for the chosen a, b
}

__global__ void __launch_bounds__(256) copy16_probe_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

}  // namespace

extern "C" {

int mbqc_probe_fp64_fma(int64_t iters, int32_t blocks, int32_t threads, double* d_out, int64_t* flops, void* stream) {
    if (iters < 1 || blocks < 1 || threads < 1 || threads > 256 || !d_out) return mbqc_set_error(MBQC_E_ARG, "mbqc_probe_fp64_fma: bad arguments");
    fp64_fma_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, 0.999999, 1e-7, d_out);
    if (flops) *flops = (int64_t)blocks * threads * iters * 16;
    return mbqc_after_launch("fp64_fma_probe_kernel");
}

int mbqc_probe_copy(const void* d_src, void* d_dst, int64_t bytes, int32_t blocks, void* stream) {
    if (!d_src || !d_dst || bytes < 16 || (bytes & 15) || blocks < 1) return mbqc_set_error(MBQC_E_ARG, "mbqc_probe_copy: bad arguments");
    copy16_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)d_src, (uint4*)d_dst, bytes / 16);
    return mbqc_after_launch("copy16_probe_kernel");
}

}  // extern "C"
