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
    const double2* target = p.target + (p.data_count > 0 ? (b << p.tab.n_out) : 0);
    double ar = 0.0, ai = 0.0;  // <t|psi>
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int d = p.tab.out_dst[i];
        if (d >= 0) {
            const double2 tg = __ldg(target + d);
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
    int64_t arow = 0, item = 0;  // angle row; data item (input / target state) of this sample
    if (live) {  // one angle per thread: a single round of loads for the whole CTA
        sample_index(p, b0 + bl, arow, item);
        double sn, cs;
        sincos_cw(__ldg(p.angles + arow * p.stride + i), sn, cs);  // one angle per thread: no table needed
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
    const double cp = sv_reg_cost<W>(p, sm, periodic, item, AngleStaged{row, 1, T, l.fixed, i, cs, ss});
    const double cm = sv_reg_cost<W>(p, sm, periodic, item, AngleStaged{row, 1, T, l.fixed, i, cs, -ss});
    p.grad[b * T + i] = (cp - cm) / (2.0 * p.shift);
    if (i == 0) {
        if (p.cost) p.cost[b] = sv_reg_cost<W>(p, sm, periodic, item, AngleStaged{row, 1, T, l.fixed, -1, 1.0, 0.0});
        if (p.status) p.status[b] = (cp == cp && cm == cm) ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM;
    }
}

// ---- prefix-sharing variant --------------------------------------------------------------------
// One thread per angle vector.  The state after measurements 0..m-1 is common to BOTH shifted
// evaluations of the parameter used at measurement m and to every later parameter, so the thread
// walks the pattern once, keeping that prefix state in (thread-private, conflict-free) shared
// memory, and forks two suffix runs per trainable measurement:
//     work = M + sum_m 2 (M - m)  ~  M^2   measurements instead of 2 T M ~ 2 M^2,
// with no phase tracking (the cost is phase invariant) and the unshifted cost for free from the
// final prefix.  Needs B >= ~100k angle vectors to fill the GPU (thread per vector).
template <int W>
__device__ __forceinline__ double fidelity_cost(const SvBatchParams& p, const double2* __restrict__ target,
                                                const double (&re)[1 << W], const double (&im)[1 << W]) {
    double ar = 0.0, ai = 0.0, n2 = 0.0;
#pragma unroll
    for (int i = 0; i < (1 << W); ++i) {
        const int d = p.tab.out_dst[i];
        if (d >= 0) {
            const double2 tg = __ldg(target + d);
            ar = fma(tg.x, re[i], fma(tg.y, im[i], ar));
            ai = fma(tg.x, im[i], fma(-tg.y, re[i], ai));
            n2 = fma(re[i], re[i], fma(im[i], im[i], n2));
        }
    }
    return 1.0 - (ar * ar + ai * ai) / n2;
}

template <int W>
__device__ __forceinline__ void rescale_state(double (&re)[1 << W], double (&im)[1 << W]) {
    double n2 = 0.0;
#pragma unroll
    for (int i = 0; i < (1 << W); ++i) n2 = fma(re[i], re[i], fma(im[i], im[i], n2));
    const double r = rsqrt(n2);
#pragma unroll
    for (int i = 0; i < (1 << W); ++i) {
        re[i] *= r;
        im[i] *= r;
    }
}

template <int W>
__global__ void __launch_bounds__(128) sv_reg_grad_prefix_kernel(const __grid_constant__ SvRegParams pp) {
    constexpr int N = 1 << W;
    extern __shared__ double2 dyn[];
    const SvBatchParams& p = pp.base;
    const PlanTables& t = p.tab;
    const int T = t.n_angles, M = t.n_steps;
    const RegSmemLayout l = reg_smem_carve(dyn, M, pp.reg.sign_pitch, pp.reg.n_fixed);
    double2* prefix = l.cs + (size_t)T * kRegThreads;  // [N][128]: element-major, thread-minor
    const int64_t b = (int64_t)blockIdx.x * kRegThreads + threadIdx.x;
    const bool live = b < p.batch;
    stage_reg_tables(pp, l);
    int64_t arow = 0, item = 0;
    if (live) {
        sample_index(p, b, arow, item);
        fetch_own_row(p.angles + arow * p.stride, l.cs + threadIdx.x, T, kRegThreads);
    }
    cp_async_wait_all();
    __syncthreads();
    if (!live) return;
    const double2* target = p.target + (p.data_count > 0 ? (item << t.n_out) : 0);
    double2* cs = l.cs + threadIdx.x;
    convert_own_row(cs, T, kRegThreads, l.trig);
    double2* pf = prefix + threadIdx.x;
    double sh_s, sh_c;
    sincos(p.shift, &sh_s, &sh_c);
    // seed -> prefix
    {
        const double2* in = p.inputs + (p.input_mode == MBQC_INPUT_BATCH ? (item << t.n_in) : 0);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t sb = (t.init_sign << (31 - i)) & 0x80000000u;
            double2 v = make_double2(t.plus_amp, 0.0);
            if (p.input_mode != MBQC_INPUT_PLUS) {
                v = __ldg(in + t.init_src[i]);
                v.x *= t.init_scale;
                v.y *= t.init_scale;
            }
            pf[i * kRegThreads] = make_double2(flip_sign(v.x, sb), flip_sign(v.y, sb));
        }
    }
    auto step_cs = [&](int m, double& c, double& s) -> int {  // returns the angle column or -1
        const int col = (int)(l.cols[m] & 0xffffu);
        if (col >= T) {
            const double2 f = l.fixed[col - T];
            c = f.x;
            s = f.y;
            return -1;
        }
        const double2 v = cs[col * kRegThreads];
        c = v.x;
        s = v.y;
        return col;
    };
    double re[N], im[N];
    bool bad = false;
    for (int m = 0; m < M; ++m) {
        double c, s;
        const int col = step_cs(m, c, s);
        if (col >= 0) {
            double cost[2];
#pragma unroll 1
            for (int sgn = 0; sgn < 2; ++sgn) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double2 v = pf[i * kRegThreads];
                    re[i] = v.x;
                    im[i] = v.y;
                }
                const double ss = sgn ? -sh_s : sh_s;
                reg_step_any<W>(re, im, (int)(l.cols[m] >> 16), c * sh_c - s * ss, s * sh_c + c * ss,
                                l.signs + m * pp.reg.sign_pitch);
                for (int m2 = m + 1; m2 < M; ++m2) {
                    double c2, s2;
                    step_cs(m2, c2, s2);
                    reg_step_any<W>(re, im, (int)(l.cols[m2] >> 16), c2, s2, l.signs + m2 * pp.reg.sign_pitch);
                    if (((m2 - m) & 15) == 15) rescale_state<W>(re, im);
                }
                cost[sgn] = fidelity_cost<W>(p, target, re, im);
            }
            p.grad[b * T + col] = (cost[0] - cost[1]) / (2.0 * p.shift);
            bad |= !(cost[0] == cost[0]) || !(cost[1] == cost[1]);
        }
        // advance the prefix by the unshifted measurement m
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double2 v = pf[i * kRegThreads];
            re[i] = v.x;
            im[i] = v.y;
        }
        reg_step_any<W>(re, im, (int)(l.cols[m] >> 16), c, s, l.signs + m * pp.reg.sign_pitch);
        if ((m & 7) == 7) rescale_state<W>(re, im);
#pragma unroll
        for (int i = 0; i < N; ++i) pf[i * kRegThreads] = make_double2(re[i], im[i]);
    }
    const double c0 = fidelity_cost<W>(p, target, re, im);
    if (p.cost) p.cost[b] = c0;
    if (p.status) p.status[b] = (bad || !(c0 == c0)) ? MBQC_STATUS_BAD_NORM : MBQC_STATUS_OK;
}

// ---- data-set average ----------------------------------------------------------------------------
// grad[p][i] = mean_s ws_grad[p*S + s][i], cost[p] = mean_s ws_cost[p*S + s], status[p] = OR_s.
// One CTA per parameter vector p; thread (lane, col) sums a strided subset of the data items in a
// fixed order (deterministic result), lanes are combined through shared memory.
constexpr int kReduceThreads = 256;

// optimiser update fused into the reduction (device-resident training loop, mbqc_train_dataset);
// formulas and operation order of mentpy/optimizers/adam.py:54-66 and sgd.py:49-59
struct OptimDev {
    int32_t kind;  // 0: none (plain gradient), MBQC_OPT_ADAM, MBQC_OPT_SGD
    int32_t nesterov;
    double step_size, b1, b2, eps, momentum;
    double bias1, bias2;  // Adam: 1 - b1^t, 1 - b2^t of this iteration
    double* x;            // [P][T] parameters, updated in place
    double* s0;           // Adam m / SGD velocity
    double* s1;           // Adam v
};

__device__ __forceinline__ void optimiser_update(const OptimDev& o, int64_t e, double g) {
    // explicit _rn operations: no FMA contraction, so the trajectory matches the host formulas
    if (o.kind == MBQC_OPT_ADAM) {
        const double m = __dadd_rn(__dmul_rn(o.b1, o.s0[e]), __dmul_rn(1.0 - o.b1, g));
        const double v = __dadd_rn(__dmul_rn(o.b2, o.s1[e]), __dmul_rn(__dmul_rn(1.0 - o.b2, g), g));
        o.s0[e] = m;
        o.s1[e] = v;
        const double m_hat = m / o.bias1, v_hat = v / o.bias2;
        o.x[e] = o.x[e] - __dmul_rn(o.step_size, m_hat) / (sqrt(v_hat) + o.eps);
    } else if (o.kind == MBQC_OPT_SGD) {
        const double vel = __dadd_rn(__dmul_rn(o.momentum, o.s0[e]), -__dmul_rn(o.step_size, g));
        o.s0[e] = vel;
        o.x[e] = o.nesterov ? __dadd_rn(__dadd_rn(o.x[e], __dmul_rn(o.momentum, vel)), -__dmul_rn(o.step_size, g))
                            : __dadd_rn(o.x[e], vel);
    }
}

__global__ void __launch_bounds__(kReduceThreads) grad_dataset_reduce_kernel(
    const double* __restrict__ ws_grad, const double* __restrict__ ws_cost, const int32_t* __restrict__ ws_status,
    int64_t S, int T, double* __restrict__ grad, double* __restrict__ cost, int32_t* __restrict__ status,
    int accumulate_status, const __grid_constant__ OptimDev opt) {
    __shared__ double part[kReduceThreads];
    __shared__ int32_t st_any;
    const int64_t p = blockIdx.x;
    const int tid = threadIdx.x;
    const int Tc = T < kReduceThreads ? T : kReduceThreads;  // columns handled per sweep
    const int L = kReduceThreads / Tc;                        // lanes per column
    const int lane = tid / Tc, c0 = tid - lane * Tc;
    const double inv = 1.0 / (double)S;
    if (tid == 0) st_any = 0;
    for (int base = 0; base < T; base += Tc) {
        const int col = base + c0;
        double sum = 0.0;
        if (lane < L && col < T) {
            const double* src = ws_grad + p * S * T + col;
            for (int64_t s = lane; s < S; s += L) sum += src[s * T];
        }
        part[tid] = sum;
        __syncthreads();
        if (lane == 0 && col < T) {
            for (int j = 1; j < L; ++j) sum += part[j * Tc + c0];
            const double g = sum * inv;
            if (grad) grad[p * T + col] = g;
            if (opt.kind) optimiser_update(opt, p * T + col, g);
        }
        __syncthreads();
    }
    double csum = 0.0;
    int32_t sany = 0;
    for (int64_t s = tid; s < S; s += kReduceThreads) {
        csum += ws_cost[p * S + s];
        sany |= ws_status[p * S + s];
    }
    part[tid] = csum;
    if (sany) atomicOr(&st_any, sany);
    __syncthreads();
    if (tid == 0) {
        for (int j = 1; j < kReduceThreads; ++j) csum += part[j];
        if (cost) cost[p] = csum * inv;
        if (status) status[p] = accumulate_status ? (status[p] | st_any) : st_any;
    }
}

}  // namespace mbqc
