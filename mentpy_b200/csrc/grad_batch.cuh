// Batched parameter-shift / central-difference gradient on top of the register SV kernel.
//
// Replaces the 2T sequential cost evaluations of mentpy/gradients/_parameter_shift.py:9-25
// (shift 1.5, divide by 2*shift -- literally a central difference, reproduced verbatim) and
// _finite_difference.py:9-25 (central) for the cost  1 - <t|rho_out|t>  with a pure target t:
// one thread per (angle vector b, parameter i) runs both shifted patterns back to back; shifted
// angle vectors are never materialised.
#pragma once
#include "sv_batch.cuh"

namespace mbqc {

template <int W>
__device__ __forceinline__ double sv_reg_cost(const SvBatchParams& p, int64_t b, int col, double sh) {
    constexpr int N = 1 << W;
    double re[N], im[N], zr, zi;
    const double n2 = sv_reg_evolve<W>(p, b, col, sh, re, im, zr, zi);
    double ar = 0.0, ai = 0.0;  // <t|psi>
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int d = p.tab.out_dst[i];
        if (d >= 0) {
            const double2 tg = __ldg(p.target + d);
            ar = fma(tg.x, re[i], fma(tg.y, im[i], ar));
            ai = fma(tg.x, im[i], fma(-tg.y, re[i], ai));
        }
    }
    return 1.0 - (ar * ar + ai * ai) / n2;
}

template <int W>
__global__ void __launch_bounds__(128) sv_reg_grad_kernel(const __grid_constant__ SvBatchParams p) {
    const int T = p.tab.n_angles;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.batch * T) return;
    // parameter index fastest: the T threads of one angle vector share its cache lines
    const int64_t b = e / T;
    const int i = (int)(e - b * T);
    const double cp = sv_reg_cost<W>(p, b, i, p.shift);
    const double cm = sv_reg_cost<W>(p, b, i, -p.shift);
    p.grad[e] = (cp - cm) / (2.0 * p.shift);
    if (p.cost && i == 0) p.cost[b] = sv_reg_cost<W>(p, b, -1, 0.0);
    if (p.status && i == 0) p.status[b] = (cp == cp && cm == cm) ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM;
}

}  // namespace mbqc
