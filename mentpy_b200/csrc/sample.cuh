// Outcome sampling with flow corrections ("force0 = False"): every measurement draws its outcome
// from the Born rule and later angles adapt to the outcomes so far.
//
// The reference's numpy simulators raise NotImplementedError for force0=False
// (np_simulator_sv.py:50-51) and Flow.adapt_angle(s) are stubs (mbqc/flow.py:105-109); the
// correction rule is the one its PennyLane circuit spells out (pennylane_simulator.py:145-153):
// outcome 1 at node i puts X on f(i) and Z on the later-measured neighbours of f(i).  A pending
// X^a Z^b on a qubit measured in the XY plane is folded into its angle, theta' = (-1)^a theta + b pi
// (cos' = (-1)^b cos, sin' = (-1)^(a+b) sin), outcome 1 adds another pi, and the output qubits get
// X^a Z^b at the end -- after which every sample's output equals the deterministic (force0) state
// up to a global phase.  Which earlier outcomes feed a step comes from the plan's feed-forward
// table (mbqc_plan_set_feedforward) as masks over a shift register of the last 32 outcomes.
//
// RNG: Philox4x32-10 (Salmon et al., SC'11), counter = (sample lo, sample hi, step, 0), key =
// seed: stateless, so a sample's outcomes do not depend on batch size, launch shape or GPU count.
#pragma once
#include "sv_batch.cuh"
#include "sv_reg.cuh"

namespace mbqc {

struct FeedForwardDev {
    uint32_t xdep, zdep, outx, outz;
};

struct SampleParams {
    const FeedForwardDev* __restrict__ ff;  // [n_steps]
    uint64_t seed;
    uint64_t sample_offset;
    int32_t outcome_mode;  // MBQC_OUTCOMES_SAMPLE | MBQC_OUTCOMES_FORCED (outcomes buffer is the input)
    int32_t correct;       // apply the byproduct X^a Z^b to the outputs
    int8_t* __restrict__ outcomes;      // [B][n_steps] (may be null when sampling)
    uint32_t* __restrict__ byproducts;  // [B]: x bits | z bits << 16, bit q = q-th output node (may be null)
    double* __restrict__ prob;          // [B] probability of the sample's outcome record (may be null)
};

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
        c[0] = n0;
        c[1] = l1;
        c[2] = n2;
        c[3] = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// uniform in [0, 1) with 53 random bits for (sample, step)
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t sample, uint32_t step) {
    uint32_t c[4] = {(uint32_t)sample, (uint32_t)(sample >> 32), step, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t bits = ((uint64_t)(c[0] >> 5) << 26) | (uint64_t)(c[1] >> 6);
    return (double)bits * 0x1.0p-53;
}

// output-index bit masks (first output node = MSB) of byproduct bits (bit q = q-th output node)
__device__ __forceinline__ uint32_t byproduct_index_mask(uint32_t bits, int k) {
    uint32_t m = 0;
    for (int q = 0; q < k; ++q) m |= ((bits >> q) & 1u) << (k - 1 - q);
    return m;
}

// sum over the pairs of slot S of |a_i + (c - i s) a_j|^2
template <int W, int S>
__device__ __forceinline__ double reg_norm0(const double (&re)[1 << W], const double (&im)[1 << W], double c, double s) {
    double n0 = 0.0;
#pragma unroll
    for (int i = 0; i < (1 << W); ++i) {
        if (i & (1 << S)) continue;
        const int j = i | (1 << S);
        const double tr = fma(c, re[j], fma(s, im[j], re[i]));
        const double ti = fma(c, im[j], fma(-s, re[j], im[i]));
        n0 = fma(tr, tr, fma(ti, ti, n0));
    }
    return n0;
}
template <int W>
__device__ __forceinline__ double reg_norm0_any(const double (&re)[1 << W], const double (&im)[1 << W], int slot,
                                                double c, double s) {
    switch (slot) {
        case 0: return reg_norm0<W, 0>(re, im, c, s);
        case 1: if constexpr (W > 1) return reg_norm0<W, 1>(re, im, c, s); break;
        case 2: if constexpr (W > 2) return reg_norm0<W, 2>(re, im, c, s); break;
        case 3: if constexpr (W > 3) return reg_norm0<W, 3>(re, im, c, s); break;
        case 4: if constexpr (W > 4) return reg_norm0<W, 4>(re, im, c, s); break;
        default: break;
    }
    return 0.0;
}

// One thread per sample, state in registers (window <= 5); out [B][2^k] complex128.
template <int W>
__global__ void __launch_bounds__(128) sv_reg_sample_kernel(const __grid_constant__ SvRegParams pp,
                                                            const __grid_constant__ SampleParams sp) {
    constexpr int N = 1 << W;
    extern __shared__ double2 dyn[];
    const SvBatchParams& p = pp.base;
    const PlanTables& t = p.tab;
    const int T = t.n_angles, M = t.n_steps;
    const RegSmemLayout l = reg_smem_carve(dyn, M, pp.reg.sign_pitch, pp.reg.n_fixed);
    const int64_t b = (int64_t)blockIdx.x * kRegThreads + threadIdx.x;
    const bool live = b < p.batch;
    stage_reg_tables(pp, l);
    if (live) fetch_own_row(p.angles + b * p.stride, l.cs + threadIdx.x, T, kRegThreads);
    cp_async_wait_all();
    __syncthreads();
    if (!live) return;
    double2* cs = l.cs + threadIdx.x;
    convert_own_row(cs, T, kRegThreads, l.trig);
    double re[N], im[N];
    {
        const double2* in = p.inputs + (p.input_mode == MBQC_INPUT_BATCH ? (b << t.n_in) : 0);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t sb = (t.init_sign << (31 - i)) & 0x80000000u;
            double2 v = make_double2(t.plus_amp, 0.0);
            if (p.input_mode != MBQC_INPUT_PLUS) {
                v = __ldg(in + t.init_src[i]);
                v.x *= t.init_scale;
                v.y *= t.init_scale;
            }
            re[i] = flip_sign(v.x, sb);
            im[i] = flip_sign(v.y, sb);
        }
    }
    double nrm = 0.0;  // squared norm of the (unnormalised) state
#pragma unroll
    for (int i = 0; i < N; ++i) nrm = fma(re[i], re[i], fma(im[i], im[i], nrm));
    const uint64_t sample = sp.sample_offset + (uint64_t)b;
    uint32_t hist = 0, bx = 0, bz = 0;
    double pb = 1.0;
    int took1 = 0;
    for (int m = 0; m < M; ++m) {
        const uint32_t cw = l.cols[m];
        const int col = (int)(cw & 0xffffu), slot = (int)(cw >> 16);
        double c, s;
        if (col >= T) {
            const double2 f = l.fixed[col - T];
            c = f.x;
            s = f.y;
        } else {
            const double2 v = cs[col * kRegThreads];
            c = v.x;
            s = v.y;
        }
        const FeedForwardDev ff = sp.ff[m];
        const uint32_t a = __popc(hist & ff.xdep) & 1u, z = __popc(hist & ff.zdep) & 1u;
        c = flip_sign(c, z << 31);
        s = flip_sign(s, (a ^ z) << 31);
        const double n0 = reg_norm0_any<W>(re, im, slot, c, s);
        const double p0 = n0 / (2.0 * nrm);
        int outcome;
        if (sp.outcome_mode == MBQC_OUTCOMES_FORCED)
            outcome = sp.outcomes[b * M + m] ? 1 : 0;
        else
            outcome = philox_uniform(sp.seed, sample, (uint32_t)m) < p0 ? 0 : 1;
        if (outcome) {  // project on (I - M)/2: angle + pi
            c = -c;
            s = -s;
        }
        // |t0|^2 + |t1|^2 = 2 |psi|^2; recompute the rare p1 << 1 case to avoid the cancellation
        double ns = outcome ? (2.0 * nrm - n0) : n0;
        if (outcome && ns < 1e-3 * nrm) ns = reg_norm0_any<W>(re, im, slot, c, s);
        pb *= outcome ? (1.0 - p0) : p0;
        reg_step_any<W>(re, im, slot, c, s, l.signs + m * pp.reg.sign_pitch);
        nrm = 2.0 * ns;  // both copies of the reduced state
        if ((m & 7) == 7) {  // keep the magnitudes bounded (the norm doubles per step at prob 1/2)
            double n2 = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) n2 = fma(re[i], re[i], fma(im[i], im[i], n2));
            const double r = rsqrt(n2);
            nrm = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                re[i] *= r;
                im[i] *= r;
                nrm = fma(re[i], re[i], fma(im[i], im[i], nrm));
            }
        }
        hist = (hist << 1) | (uint32_t)outcome;
        if (outcome) {
            bx ^= ff.outx;
            bz ^= ff.outz;
        }
        took1 |= outcome;
        if (sp.outcome_mode != MBQC_OUTCOMES_FORCED && sp.outcomes) sp.outcomes[b * M + m] = (int8_t)outcome;
    }
    const int k = t.n_out;
    double n2 = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (t.out_dst[i] >= 0) n2 = fma(re[i], re[i], fma(im[i], im[i], n2));
    const bool ok = (n2 > 0.0) && isfinite(n2) && (pb == pb);
    if (p.status) p.status[b] = (ok ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM) | (took1 ? MBQC_STATUS_OUTCOME1 : 0);
    if (sp.byproducts) sp.byproducts[b] = bx | (bz << 16);
    if (sp.prob) sp.prob[b] = pb;
    const uint32_t xm = sp.correct ? byproduct_index_mask(bx, k) : 0u;
    const uint32_t zm = sp.correct ? byproduct_index_mask(bz, k) : 0u;
    const double r = rsqrt(n2);
    double2* o = p.out + (b << k);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int d = t.out_dst[i];
        if (d >= 0) {  // Z^z then X^x on the output register
            const uint32_t sg = (uint32_t)(__popc((uint32_t)d & zm) & 1) << 31;
            o[(uint32_t)d ^ xm] = make_double2(flip_sign(re[i] * r, sg), flip_sign(im[i] * r, sg));
        }
    }
}

// Shared-memory variant for 6 <= window <= 12: a group of 2^tps_log2 threads per shot, amplitudes in
// shared memory (the layout of sv_smem_kernel); every thread of the group evaluates the same Born
// probability (group reduction) and the same Philox draw, so the outcome is group-uniform.
__global__ void sv_smem_sample_kernel(const __grid_constant__ SvBatchParams p, const __grid_constant__ SampleParams sp,
                                      int tps_log2, int spb) {
    extern __shared__ double2 smem[];
    __shared__ double red[32];
    const PlanTables& t = p.tab;
    const int w = t.window, M = t.n_steps;
    const int tps = 1 << tps_log2;
    const int ls = threadIdx.x >> tps_log2;
    const int tid = threadIdx.x & (tps - 1);
    const int64_t b = (int64_t)blockIdx.x * spb + ls;
    const bool live = b < p.batch;
    const int64_t be = live ? b : 0;
    double2* psi = smem + ((size_t)ls << w);
    double2* cs = smem + ((size_t)spb << w) + ((size_t)ls << tps_log2);  // staged (cos, sin), as in sv_smem_kernel
    const uint64_t n = 1ull << w, half = n >> 1;
    double nrm = 0.0;
    {
        const double2* in = (p.input_mode == MBQC_INPUT_PLUS)
                                ? nullptr
                                : p.inputs + (p.input_mode == MBQC_INPUT_BATCH ? (be << t.n_in) : 0);
        for (uint64_t i = tid; i < n; i += tps) {
            double2 v = make_double2(t.plus_amp, 0.0);
            if (in) {
                v = __ldg(in + init_source_index(t, i));
                v.x *= t.init_scale;
                v.y *= t.init_scale;
            }
            if (init_sign_bit(t, i)) {
                v.x = -v.x;
                v.y = -v.y;
            }
            psi[i] = v;
            nrm = fma(v.x, v.x, fma(v.y, v.y, nrm));
        }
    }
    nrm = group_sum(nrm, tps_log2, ls, red);  // contains the barrier that publishes psi for tps > 32
    __syncthreads();
    const double* row = p.angles + be * p.stride;
    const uint64_t sample = sp.sample_offset + (uint64_t)be;
    uint32_t hist = 0, bx = 0, bz = 0;
    double pb = 1.0;
    int took1 = 0;
    for (int m = 0; m < M; ++m) {
        const StepDev st = p.steps[m];
        if ((m & (tps - 1)) == 0) {
            stage_group_angles(p, row, m, tid, cs);
            group_barrier(tps_log2);
        }
        double c = cs[m & (tps - 1)].x, s = cs[m & (tps - 1)].y;
        const FeedForwardDev ff = sp.ff[m];
        const uint32_t a = __popc(hist & ff.xdep) & 1u, z = __popc(hist & ff.zdep) & 1u;
        c = flip_sign(c, z << 31);
        s = flip_sign(s, (a ^ z) << 31);
        const uint64_t bit = 1ull << st.slot;
        double n0 = 0.0;  // || <0_theta| psi ||^2 (times 2): Born weight of outcome 0
        for (uint64_t g = tid; g < half; g += tps) {
            const uint64_t i0 = insert_zero(g, st.slot);
            const double2 x = psi[i0], y = psi[i0 | bit];
            const double tr = fma(c, y.x, fma(s, y.y, x.x)), ti = fma(c, y.y, fma(-s, y.x, x.y));
            n0 = fma(tr, tr, fma(ti, ti, n0));
        }
        n0 = group_sum(n0, tps_log2, ls, red);
        const double p0 = n0 / (2.0 * nrm);
        int outcome;
        if (sp.outcome_mode == MBQC_OUTCOMES_FORCED)
            outcome = sp.outcomes[be * M + m] ? 1 : 0;
        else
            outcome = philox_uniform(sp.seed, sample, (uint32_t)m) < p0 ? 0 : 1;
        if (outcome) {  // project on (I - M)/2: angle + pi
            c = -c;
            s = -s;
        }
        pb *= outcome ? (1.0 - p0) : p0;
        double ns = 0.0;  // squared norm of the projected state, summed while it is written
        for (uint64_t g = tid; g < half; g += tps) {
            const uint64_t i0 = insert_zero(g, st.slot);
            const double2 x = psi[i0], y = psi[i0 | bit];
            double2 tt;
            tt.x = fma(c, y.x, fma(s, y.y, x.x));
            tt.y = fma(c, y.y, fma(-s, y.x, x.y));
            ns = fma(tt.x, tt.x, fma(tt.y, tt.y, ns));
            psi[i0] = tt;
            if (parity64(i0 & st.nbr_mask)) {
                tt.x = -tt.x;
                tt.y = -tt.y;
            }
            psi[i0 | bit] = tt;
        }
        ns = group_sum(ns, tps_log2, ls, red);
        nrm = 2.0 * ns;  // both copies of the reduced state
        __syncthreads();
        if ((m & 7) == 7) {  // keep the magnitudes bounded (the norm doubles per step at prob 1/2)
            const double r = rsqrt(nrm);
            for (uint64_t i = tid; i < n; i += tps) {
                psi[i].x *= r;
                psi[i].y *= r;
            }
            nrm = 1.0;
            __syncthreads();
        }
        hist = (hist << 1) | (uint32_t)outcome;
        if (outcome) {
            bx ^= ff.outx;
            bz ^= ff.outz;
        }
        took1 |= outcome;
        if (live && tid == 0 && sp.outcome_mode != MBQC_OUTCOMES_FORCED && sp.outcomes) sp.outcomes[b * M + m] = (int8_t)outcome;
    }
    const int k = t.n_out;
    const uint32_t no = 1u << k;
    double n2 = 0.0;
    for (uint32_t o = tid; o < no; o += tps) {
        const double2 v = psi[output_state_index(t, o)];
        n2 = fma(v.x, v.x, fma(v.y, v.y, n2));
    }
    n2 = group_sum(n2, tps_log2, ls, red);
    if (!live) return;
    const bool ok = (n2 > 0.0) && isfinite(n2) && (pb == pb);
    if (tid == 0) {
        if (p.status) p.status[b] = (ok ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM) | (took1 ? MBQC_STATUS_OUTCOME1 : 0);
        if (sp.byproducts) sp.byproducts[b] = bx | (bz << 16);
        if (sp.prob) sp.prob[b] = pb;
    }
    const uint32_t xm = sp.correct ? byproduct_index_mask(bx, k) : 0u;
    const uint32_t zm = sp.correct ? byproduct_index_mask(bz, k) : 0u;
    const double r = rsqrt(n2);
    double2* o = p.out + (b << k);
    for (uint32_t d = tid; d < no; d += tps) {  // Z^z then X^x on the output register
        const double2 v = psi[output_state_index(t, d)];
        const uint32_t sg = (uint32_t)(__popc(d & zm) & 1) << 31;
        o[d ^ xm] = make_double2(flip_sign(v.x * r, sg), flip_sign(v.y * r, sg));
    }
}

}  // namespace mbqc
