// State helpers of the reference's calculator module as kernels (mentpy/calculator/state_ops.py):
//   pure "partial trace"  : SUM over the traced qubits, then renormalise      (:42-74)
//   mixed partial trace   : sigma[r,c] = sum_m rho[(r,m),(c,m)]               (:77-119)
//   pure2density          : |psi><psi|                                         (:16-39)
// Qubit 0 is the most significant bit (reference convention).  Small, single-state utilities: one
// CTA, block reductions by warp shuffle.
#pragma once
#include "common.cuh"

namespace mbqc {

// deposit the low bits of `v` into the set bits of `mask` (ascending)
__device__ __forceinline__ uint32_t deposit_bits(uint32_t v, uint32_t mask) {
    uint32_t out = 0;
    for (uint32_t m = mask; m; m &= m - 1) {
        const uint32_t bit = m & (0u - m);
        if (v & 1u) out |= bit;
        v >>= 1;
    }
    return out;
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += red[k];
    return tot;
}

// keep_mask / trace_mask: bit masks over the n index bits (bit n-1-q <-> qubit q)
__global__ void trace_pure_kernel(const double2* __restrict__ psi, double2* __restrict__ out, int n,
                                  uint32_t keep_mask, uint32_t trace_mask) {
    __shared__ double red[32];
    const int nk = __popc(keep_mask), nt = __popc(trace_mask);
    double n2 = 0.0;
    for (uint32_t r = threadIdx.x; r < (1u << nk); r += blockDim.x) {
        const uint32_t base = deposit_bits(r, keep_mask);
        double2 acc = make_double2(0.0, 0.0);
        for (uint32_t m = 0; m < (1u << nt); ++m) {
            const double2 v = psi[base | deposit_bits(m, trace_mask)];
            acc.x += v.x;
            acc.y += v.y;
        }
        out[r] = acc;
        n2 += acc.x * acc.x + acc.y * acc.y;
    }
    n2 = block_sum(n2, red);
    const double inv = rsqrt(n2);
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < (1u << nk); r += blockDim.x) {
        out[r].x *= inv;
        out[r].y *= inv;
    }
}

__global__ void trace_mixed_kernel(const double2* __restrict__ rho, double2* __restrict__ out, int n,
                                   uint32_t keep_mask, uint32_t trace_mask) {
    const int nk = __popc(keep_mask), nt = __popc(trace_mask);
    const uint32_t dim_out = 1u << nk;
    const uint64_t total = (uint64_t)dim_out * dim_out;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(e >> nk), c = (uint32_t)(e & (dim_out - 1));
        const uint32_t rb = deposit_bits(r, keep_mask), cb = deposit_bits(c, keep_mask);
        double2 acc = make_double2(0.0, 0.0);
        for (uint32_t m = 0; m < (1u << nt); ++m) {
            const uint32_t mb = deposit_bits(m, trace_mask);
            const double2 v = rho[((uint64_t)(rb | mb) << n) | (cb | mb)];
            acc.x += v.x;
            acc.y += v.y;
        }
        out[e] = acc;
    }
}

__global__ void pure2density_kernel(const double2* __restrict__ psi, double2* __restrict__ out, int n) {
    const uint64_t total = 1ull << (2 * n);
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const double2 x = psi[e >> n], y = psi[e & ((1ull << n) - 1)];
        out[e] = make_double2(x.x * y.x + x.y * y.y, x.y * y.x - x.x * y.y);
    }
}

}  // namespace mbqc
