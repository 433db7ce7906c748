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

template <int W, class AngleSrc>
__device__ __forceinline__ double sv_reg_cost(const SvBatchParams& p, const StepDev* steps, int64_t b,
                                              const AngleSrc& ang) {
    constexpr int N = 1 << W;
    double re[N], im[N], zr, zi;
    const double n2 = sv_reg_evolve<W>(p, steps, b, ang, re, im, zr, zi);
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

// One thread per (angle vector b, parameter i), parameter index fastest, so the T threads of one
// angle vector sit next to each other and share its staged (cos, sin) row: one sincos per angle
// instead of one per angle per shifted evaluation.  The CTA covers `spb` whole angle vectors.
template <int W>
__global__ void __launch_bounds__(128) sv_reg_grad_kernel(const __grid_constant__ SvBatchParams p, int tp, int spb) {
    extern __shared__ double2 dyn[];
    StepDev* s_steps = reinterpret_cast<StepDev*>(dyn);
    double2* s_cs = dyn + 3 * p.tab.n_steps;
    const int T = p.tab.n_angles;
    const int64_t b0 = (int64_t)blockIdx.x * spb;
    const int samples = (int)min((int64_t)spb, p.batch - b0);
    stage_plan_and_angles(p, s_steps, s_cs, tp, b0, samples, tp > 0);
    const int bl = threadIdx.x / T;
    const int i = threadIdx.x - bl * T;
    if (bl >= samples) return;
    const int64_t b = b0 + bl;
    double cp, cm, c0 = 0.0;
    if (tp > 0) {
        double ss, cs;
        sincos(p.shift, &ss, &cs);
        cp = sv_reg_cost<W>(p, s_steps, b, AngleStaged{s_cs + bl * tp, i, cs, ss});
        cm = sv_reg_cost<W>(p, s_steps, b, AngleStaged{s_cs + bl * tp, i, cs, -ss});
        if (p.cost && i == 0) c0 = sv_reg_cost<W>(p, s_steps, b, AngleStaged{s_cs + bl * tp, -1, 1.0, 0.0});
    } else {
        const double* row = p.angles + b * p.stride;
        cp = sv_reg_cost<W>(p, s_steps, b, AngleGlobal{row, i, p.shift});
        cm = sv_reg_cost<W>(p, s_steps, b, AngleGlobal{row, i, -p.shift});
        if (p.cost && i == 0) c0 = sv_reg_cost<W>(p, s_steps, b, AngleGlobal{row, -1, 0.0});
    }
    p.grad[b * T + i] = (cp - cm) / (2.0 * p.shift);
    if (p.cost && i == 0) p.cost[b] = c0;
    if (p.status && i == 0) p.status[b] = (cp == cp && cm == cm) ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM;
}

}  // namespace mbqc
