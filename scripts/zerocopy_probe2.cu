// Probe 2: H2D by DMA (chunked) + kernel writing results straight into page-locked host memory.
#include <cuda_runtime.h>
#include <cstdio>
#include <chrono>
__global__ void work(const double* __restrict__ a, double2* __restrict__ o, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0;
    for (int j = 0; j < 10; ++j) s += a[i * 10 + j];
    // staged, coalesced 16-byte stores: 4 consecutive double2 per sample
    for (int j = 0; j < 4; ++j) o[i * 4 + j] = make_double2(s, j);
}
// variant: smem-staged so that a warp writes 512 contiguous bytes per instruction
__global__ void work_coalesced(const double* __restrict__ a, double2* __restrict__ o, long n) {
    __shared__ double2 st[128 * 4];
    long b0 = blockIdx.x * (long)blockDim.x; long i = b0 + threadIdx.x;
    double s = 0;
    if (i < n) for (int j = 0; j < 10; ++j) s += a[i * 10 + j];
    for (int j = 0; j < 4; ++j) st[threadIdx.x * 4 + j] = make_double2(s, j);
    __syncthreads();
    long cnt = min((long)blockDim.x, n - b0) * 4;
    for (long e = threadIdx.x; e < cnt; e += blockDim.x) o[b0 * 4 + e] = st[e];
}
int main() {
    const long B = 65536; const int T = 10, K = 4, S = 4;
    double* h_in; double2* h_out; cudaHostAlloc(&h_in, B * T * 8, cudaHostAllocMapped); cudaHostAlloc(&h_out, B * K * 16, cudaHostAllocMapped);
    double* d_in; cudaMalloc(&d_in, B * T * 8);
    cudaStream_t st[S]; for (int i = 0; i < S; ++i) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    for (int variant = 0; variant < 2; ++variant)
    for (int chunks : {1, 2, 4, 8}) {
        double best = 1e9;
        for (int rep = 0; rep < 20; ++rep) {
            cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            long per = B / chunks;
            for (int c = 0; c < chunks; ++c) {
                cudaStream_t s = st[c % S]; long lo = c * per;
                cudaMemcpyAsync(d_in + lo * T, h_in + lo * T, per * T * 8, cudaMemcpyHostToDevice, s);
                if (variant == 0) work<<<(per + 127) / 128, 128, 0, s>>>(d_in + lo * T, h_out + lo * K, per);
                else work_coalesced<<<(per + 127) / 128, 128, 0, s>>>(d_in + lo * T, h_out + lo * K, per);
            }
            for (int i = 0; i < S; ++i) cudaStreamSynchronize(st[i]);
            double tot = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            if (tot < best) best = tot;
        }
        printf("variant=%d (H2D DMA + kernel writes host) chunks=%d: best %.1f us\n", variant, chunks, best);
    }
    return 0;
}
