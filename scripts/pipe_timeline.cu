// Standalone probe: timeline of a chunked H2D -> kernel -> D2H pipeline on non-blocking streams.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <chrono>
__global__ void dummy(const double* a, double2* o, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) { double s = 0; for (int j = 0; j < 10; ++j) s += a[i * 10 + j]; for (int j = 0; j < 4; ++j) o[i * 4 + j] = make_double2(s, j); }
}
int main() {
    const long B = 65536; const int T = 10, K = 4, S = 4;
    double* h_in; double2* h_out; cudaHostAlloc(&h_in, B * T * 8, 0); cudaHostAlloc(&h_out, B * K * 16, 0);
    double* d_in; double2* d_out; cudaMalloc(&d_in, B * T * 8); cudaMalloc(&d_out, B * K * 16);
    cudaStream_t st[S]; for (int i = 0; i < S; ++i) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    for (int chunks : {1, 2, 4, 8}) {
        std::vector<cudaEvent_t> ev(chunks * 3 + 1);
        for (auto& e : ev) cudaEventCreate(&e);
        for (int rep = 0; rep < 3; ++rep) {
            cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            cudaEventRecord(ev[0], st[0]);
            for (int i = 1; i < S; ++i) cudaStreamWaitEvent(st[i], ev[0], 0);
            long per = B / chunks;
            std::vector<double> cpu_t;
            for (int c = 0; c < chunks; ++c) {
                cudaStream_t s = st[c % S]; long lo = c * per;
                cudaMemcpyAsync(d_in + lo * T, h_in + lo * T, per * T * 8, cudaMemcpyHostToDevice, s);
                cudaEventRecord(ev[1 + 3 * c], s);
                dummy<<<(per + 127) / 128, 128, 0, s>>>(d_in + lo * T, d_out + lo * K, per);
                cudaEventRecord(ev[2 + 3 * c], s);
                cudaMemcpyAsync(h_out + lo * K, d_out + lo * K, per * K * 16, cudaMemcpyDeviceToHost, s);
                cudaEventRecord(ev[3 + 3 * c], s);
                cpu_t.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
            }
            for (int i = 0; i < S; ++i) cudaStreamSynchronize(st[i]);
            double tot = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            if (rep == 2) {
                printf("chunks=%d total(host) %.1f us\n", chunks, tot);
                for (int c = 0; c < chunks; ++c) {
                    float a, b, d; cudaEventElapsedTime(&a, ev[0], ev[1 + 3 * c]); cudaEventElapsedTime(&b, ev[0], ev[2 + 3 * c]); cudaEventElapsedTime(&d, ev[0], ev[3 + 3 * c]);
                    printf("  chunk %d: h2d done %.1f  kernel done %.1f  d2h done %.1f   (cpu issued by %.1f us)\n", c, a * 1e3, b * 1e3, d * 1e3, cpu_t[c]);
                }
            }
        }
    }
    return 0;
}
