// Probe: dedicated H2D / compute / D2H streams with events vs per-chunk streams.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <chrono>
__global__ void dummy(const double* a, double2* o, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) { double s = 0; for (int j = 0; j < 10; ++j) s += a[i * 10 + j]; for (int j = 0; j < 4; ++j) o[i * 4 + j] = make_double2(s, j); }
}
int main() {
    const long B = 65536; const int T = 10, K = 4;
    double* h_in; double2* h_out; cudaHostAlloc(&h_in, B * T * 8, 0); cudaHostAlloc(&h_out, B * K * 16, 0);
    double* d_in; double2* d_out; cudaMalloc(&d_in, B * T * 8); cudaMalloc(&d_out, B * K * 16);
    cudaStream_t sin_, sk, sout; cudaStreamCreateWithFlags(&sin_, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sout, cudaStreamNonBlocking);
    for (int chunks : {1, 2, 3, 4, 6, 8}) {
        std::vector<cudaEvent_t> e_in(chunks), e_k(chunks);
        for (auto& e : e_in) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        for (auto& e : e_k) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        double best = 1e9;
        for (int rep = 0; rep < 30; ++rep) {
            cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            long per = B / chunks;
            for (int c = 0; c < chunks; ++c) {
                long lo = c * per, n = (c == chunks - 1) ? B - lo : per;
                cudaMemcpyAsync(d_in + lo * T, h_in + lo * T, n * T * 8, cudaMemcpyHostToDevice, sin_);
                cudaEventRecord(e_in[c], sin_);
                cudaStreamWaitEvent(sk, e_in[c], 0);
                dummy<<<(n + 127) / 128, 128, 0, sk>>>(d_in + lo * T, d_out + lo * K, n);
                cudaEventRecord(e_k[c], sk);
                cudaStreamWaitEvent(sout, e_k[c], 0);
                cudaMemcpyAsync(h_out + lo * K, d_out + lo * K, n * K * 16, cudaMemcpyDeviceToHost, sout);
            }
            cudaStreamSynchronize(sout);
            double tot = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            if (tot < best) best = tot;
        }
        printf("3-stream pipeline chunks=%d: best %.1f us\n", chunks, best);
    }
    return 0;
}
