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
// instead of one per angle per shifted evaluation, and a shift is a rotation of that pair.
// The CTA covers `spb` whole angle vectors (spb * T <= 128 threads).
template <int W>
__global__ void __launch_bounds__(128) sv_reg_grad_kernel(const __grid_constant__ SvBatchParams p, int staged, int spb) {
    extern __shared__ double2 dyn[];
    const int T = p.tab.n_angles;
    StepDev* s_steps = reinterpret_cast<StepDev*>(dyn);
    double2* s_cs = dyn + 3 * p.tab.n_steps;  // [spb][T] : element e = bl * T + i, pitch 1 per column
    const int64_t b0 = (int64_t)blockIdx.x * spb;
    const int samples = (int)min((int64_t)spb, p.batch - b0);
    stage_steps(p, s_steps);
    const int bl = threadIdx.x / T;
    const int i = threadIdx.x - bl * T;
    const bool live = bl < samples;
    if (staged && live) {  // one angle per thread: a single round of loads for the whole CTA
        double sn, cs;
        sincos(__ldg(p.angles + (b0 + bl) * p.stride + i), &sn, &cs);
        s_cs[threadIdx.x] = make_double2(cs, sn);
    }
    cp_async_wait_all();
    __syncthreads();
    if (!live) return;
    const int64_t b = b0 + bl;
    double cp, cm, c0 = 0.0;
    if (staged) {
        double ss, cs;
        sincos(p.shift, &ss, &cs);
        const double2* row = s_cs + bl * T;
        cp = sv_reg_cost<W>(p, s_steps, b, AngleStaged{row, 1, i, cs, ss});
        cm = sv_reg_cost<W>(p, s_steps, b, AngleStaged{row, 1, i, cs, -ss});
        if (p.cost && i == 0) c0 = sv_reg_cost<W>(p, s_steps, b, AngleStaged{row, 1, -1, 1.0, 0.0});
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
