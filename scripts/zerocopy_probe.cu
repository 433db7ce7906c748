// Probe: kernel reading/writing page-locked HOST memory directly (UVA zero-copy) vs cudaMemcpyAsync.
#include <cuda_runtime.h>
#include <cstdio>
#include <chrono>
// each CTA: 128 threads; reads a contiguous tile of IN_PER doubles per thread-row, writes tile out
template <int MODE>  // 0: ld.global, 1: cp.async to smem
__global__ void zc(const double2* __restrict__ in, double2* __restrict__ out, long n_in16, long n_out16) {
    // in: n_in16 16-byte elements, out: n_out16 16-byte elements; grid-stride coalesced
    extern __shared__ double2 sm[];
    const long tiles = gridDim.x;
    const long in_per = n_in16 / tiles, out_per = n_out16 / tiles;
    const double2* src = in + blockIdx.x * in_per;
    double2* dst = out + blockIdx.x * out_per;
    double acc = 0;
    if (MODE == 0) {
        for (long i = threadIdx.x; i < in_per; i += blockDim.x) { double2 v = src[i]; sm[i] = v; }
    } else {
        for (long i = threadIdx.x; i < in_per; i += blockDim.x) {
            unsigned d = (unsigned)__cvta_generic_to_shared(sm + i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + i));
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    }
    __syncthreads();
    for (long i = threadIdx.x; i < in_per; i += blockDim.x) acc += sm[i].x;
    for (long i = threadIdx.x; i < out_per; i += blockDim.x) dst[i] = make_double2(acc, (double)i);
}
int main() {
    const long B = 65536; const long n_in16 = B * 10 / 2, n_out16 = B * 4;
    double2 *h_in, *h_out; cudaHostAlloc(&h_in, n_in16 * 16, cudaHostAllocMapped); cudaHostAlloc(&h_out, n_out16 * 16, cudaHostAllocMapped);
    for (long i = 0; i < n_in16; ++i) h_in[i] = make_double2(1.0, 2.0);
    for (int blocks : {128, 256, 512, 1024, 2048}) {
        for (int mode = 0; mode < 2; ++mode) {
            size_t smem = (n_in16 / blocks) * 16;
            auto run = [&]() { if (mode == 0) zc<0><<<blocks, 128, smem>>>(h_in, h_out, n_in16, n_out16); else zc<1><<<blocks, 128, smem>>>(h_in, h_out, n_in16, n_out16); };
            for (int i = 0; i < 3; ++i) run();
            cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            const int N = 20;
            for (int i = 0; i < N; ++i) run();
            cudaError_t e = cudaDeviceSynchronize();
            double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / N;
            printf("blocks=%4d mode=%d: %.1f us per pass (%.1f GB/s in, %.1f GB/s out) %s\n", blocks, mode, us, n_in16 * 16 / us / 1e3, n_out16 * 16 / us / 1e3, cudaGetErrorString(e));
        }
    }
    return 0;
}
