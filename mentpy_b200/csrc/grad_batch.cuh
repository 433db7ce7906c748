// Batched parameter-shift / central-difference gradient on top of the register SV kernel.
//
// Replaces the 2T sequential cost evaluations of mentpy/gradients/_parameter_shift.py:9-25
// (shift 1.5, divide by 2*shift -- literally a central difference, reproduced verbatim) and
// _finite_difference.py:9-25 (central) for the cost  1 - <t|rho_out|t>  with a pure target t.
// One thread per (angle vector b, parameter i), parameter index fastest, so the T threads of one
// angle vector sit next to each other and share its staged (cos, sin) row: one sincos per angle
// instead of one per angle per shifted evaluation, and a shift is a rotation of that pair.
// Shifted angle vectors are never materialised.  The CTA covers `spb` whole angle vectors.
#pragma once
#include "sv_reg.cuh"

namespace mbqc {

template <int W, class AngleSrc>
__device__ __forceinline__ double sv_reg_cost(const SvBatchParams& p, const RegSmem& sm, bool periodic,
                                              int64_t b, const AngleSrc& ang) {
    constexpr int N = 1 << W;
    double re[N], im[N], zr, zi;
    const double n2 = sv_reg_evolve<W, true, true>(p, sm, periodic, b, ang, re, im, zr, zi);
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
__global__ void __launch_bounds__(128) sv_reg_grad_kernel(const __grid_constant__ SvRegParams pp, int spb) {
    extern __shared__ double2 dyn[];
    const SvBatchParams& p = pp.base;
    const int T = p.tab.n_angles, M = p.tab.n_steps;
    // cs tile: [spb][T] row-major (pitch 1 per column inside a row)
    const RegSmemLayout l = reg_smem_carve(dyn, M, pp.reg.sign_pitch, pp.reg.n_fixed);
    const int64_t b0 = (int64_t)blockIdx.x * spb;
    const int samples = (int)min((int64_t)spb, p.batch - b0);
    stage_reg_tables(pp, l);
    const int bl = threadIdx.x / T;
    const int i = threadIdx.x - bl * T;
    const bool live = bl < samples;
    if (live) {  // one angle per thread: a single round of loads for the whole CTA
        double sn, cs;
        sincos_cw(__ldg(p.angles + (b0 + bl) * p.stride + i), sn, cs);  // one angle per thread: no table needed
        l.cs[threadIdx.x] = make_double2(cs, sn);
    }
    cp_async_wait_all();
    __syncthreads();
    if (!live) return;
    const RegSmem sm{l.cols, l.signs, pp.reg.sign_pitch};
    const bool periodic = pp.reg.periodic != 0;
    const int64_t b = b0 + bl;
    double ss, cs;
    sincos(p.shift, &ss, &cs);
    const double2* row = l.cs + bl * T;
    const double cp = sv_reg_cost<W>(p, sm, periodic, b, AngleStaged{row, 1, T, l.fixed, i, cs, ss});
    const double cm = sv_reg_cost<W>(p, sm, periodic, b, AngleStaged{row, 1, T, l.fixed, i, cs, -ss});
    p.grad[b * T + i] = (cp - cm) / (2.0 * p.shift);
    if (i == 0) {
        if (p.cost) p.cost[b] = sv_reg_cost<W>(p, sm, periodic, b, AngleStaged{row, 1, T, l.fixed, -1, 1.0, 0.0});
        if (p.status) p.status[b] = (cp == cp && cm == cm) ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM;
    }
}

}  // namespace mbqc
