// Register-resident batched density-matrix kernel (window w <= 4): one LANE per row of rho.
//
// A sample is handled by 2^w lanes of a warp (2 samples per warp at w = 4); lane r keeps row r of
// rho (2^w complex numbers) in registers.  For the measurement of slot s the column bit is a
// register index (compile-time pairs) and the row bit is a lane bit (one __shfl_xor exchange):
//
//     u_a    = q_{a0} rho_{a0} + q_{a1} rho_{a1}        per lane, a = its row bit     (registers)
//     sigma  = u_0 + u_1                                 exchange with lane ^ (1 << s) (shuffles)
//     prob   = sum of Re sigma on the diagonal groups    2^w-lane shuffle reduction
//     rho'_{ab} = sigma / (2 prob) * sign(r,a) sign(c,b) back into the same registers
//
// No shared-memory traffic, no barriers and no index arithmetic inside the step loop: ~1/3 of the
// instructions of dm_smem_kernel (which stays for w = 5, 6).  Same arithmetic and quirks
// (np_simulator_dm.py:151-346: outcome 1 iff prob0 < 1e-4, per-step normalisation, NaN status),
// optional channel folded into the projector coefficients, channel on the output qubits and the
// output gather done once at the end through a small shared-memory stage.
#pragma once
#include "dm_batch.cuh"

namespace mbqc {

template <int W, int S>
__device__ __forceinline__ void dm_reg_stage(double (&re)[1 << W], double (&im)[1 << W], const MeasCoef& q,
                                             uint32_t r, uint64_t nbr_mask, int lps_base, bool append,
                                             int& outcome, int& bad) {
    constexpr int N = 1 << W;
    constexpr int NP = N >> 1;
    const uint32_t a = (r >> S) & 1u;
    // u = alpha * rho_{a0} + beta * rho_{a1}; a = 0: (q00, q01), a = 1: (conj(q01), q11)
    const double ar = a ? q.q01r : q.q00, ai = a ? -q.q01i : 0.0;
    const double br = a ? q.q11 : q.q01r, bi = a ? 0.0 : q.q01i;
    double sr[NP], si[NP];
    const uint32_t r0 = r & ~(1u << S);
    double tr0 = 0.0;
    int p = 0;
#pragma unroll
    for (int c0 = 0; c0 < N; ++c0) {
        if (c0 & (1 << S)) continue;
        const int c1 = c0 | (1 << S);
        double ur = fma(ar, re[c0], fma(-ai, im[c0], fma(br, re[c1], -bi * im[c1])));
        double ui = fma(ar, im[c0], fma(ai, re[c0], fma(br, im[c1], bi * re[c1])));
        ur += __shfl_xor_sync(0xffffffffu, ur, 1 << S);
        ui += __shfl_xor_sync(0xffffffffu, ui, 1 << S);
        sr[p] = ur;
        si[p] = ui;
        if (a == 0 && (uint32_t)c0 == r0) tr0 += ur;  // diagonal group, counted once
        ++p;
    }
#pragma unroll
    for (int o = (1 << W) >> 1; o > 0; o >>= 1) tr0 += __shfl_xor_sync(0xffffffffu, tr0, o);
    outcome = (tr0 < 1e-4) ? 1 : 0;  // np_simulator_dm.py:335-338
    double prob = tr0;
    if (__any_sync(0xffffffffu, outcome)) {
        // rare: sigma1 = tr_s(rho) - sigma0, tr_s(rho) = rho_00 + rho_11 (own diagonal block + partner's)
        double trf = 0.0;
        p = 0;
#pragma unroll
        for (int c0 = 0; c0 < N; ++c0) {
            if (c0 & (1 << S)) continue;
            const int c1 = c0 | (1 << S);
            double fr = a ? re[c1] : re[c0], fi = a ? im[c1] : im[c0];
            fr += __shfl_xor_sync(0xffffffffu, fr, 1 << S);
            fi += __shfl_xor_sync(0xffffffffu, fi, 1 << S);
            if (a == 0 && (uint32_t)c0 == r0) trf += fr;
            if (outcome) {
                sr[p] = fr - sr[p];
                si[p] = fi - si[p];
            }
            ++p;
        }
#pragma unroll
        for (int o = (1 << W) >> 1; o > 0; o >>= 1) trf += __shfl_xor_sync(0xffffffffu, trf, o);
        if (outcome) prob = trf - tr0;
    }
    if (!(prob > 0.0) || !isfinite(prob)) bad = 1;
    const double sc = 0.5 / prob;
    const uint32_t pr = (a & (append ? parity64((uint64_t)r0 & nbr_mask) : 0u)) << 31;
    p = 0;
#pragma unroll
    for (int c0 = 0; c0 < N; ++c0) {
        if (c0 & (1 << S)) continue;
        const int c1 = c0 | (1 << S);
        const double vr = sr[p] * sc, vi = si[p] * sc;
        const uint32_t pc = append ? (parity64((uint64_t)c0 & nbr_mask) << 31) : 0u;
        re[c0] = flip_sign(vr, pr);
        im[c0] = flip_sign(vi, pr);
        re[c1] = flip_sign(vr, pr ^ pc);
        im[c1] = flip_sign(vi, pr ^ pc);
        ++p;
    }
    (void)lps_base;
}

template <int W>
__global__ void __launch_bounds__(128) dm_reg_kernel(const __grid_constant__ DmBatchParams p) {
    constexpr int N = 1 << W;           // lanes per sample = columns per lane
    constexpr int SPW = 32 / N;         // samples per warp
    constexpr int SPB = 4 * SPW;        // samples per CTA (4 warps)
    extern __shared__ double2 stage[];  // [SPB][N][N] rows of rho for the output phase
    const PlanTables& t = p.tab;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / N;
    const uint32_t r = lane % N;
    const int ls = warp * SPW + grp;
    const int64_t b = (int64_t)blockIdx.x * SPB + ls;
    const bool live = b < p.batch;
    const int64_t be = live ? b : 0;
    const int lane_base = grp * N;

    // seed: psi[r] on lane r, rho[r][c] = psi[r] conj(psi[c])
    double2 psi;
    {
        const double2* in = (p.input_mode == MBQC_INPUT_PLUS)
                                ? nullptr
                                : p.inputs + (p.input_mode == MBQC_INPUT_BATCH ? (be << t.n_in) : 0);
        psi = make_double2(t.plus_amp, 0.0);
        if (in) {
            psi = __ldg(in + init_source_index(t, r));
            psi.x *= t.init_scale;
            psi.y *= t.init_scale;
        }
        if (init_sign_bit(t, r)) {
            psi.x = -psi.x;
            psi.y = -psi.y;
        }
    }
    double re[N], im[N];
#pragma unroll
    for (int c = 0; c < N; ++c) {
        const double yr = __shfl_sync(0xffffffffu, psi.x, lane_base + c);
        const double yi = __shfl_sync(0xffffffffu, psi.y, lane_base + c);
        re[c] = psi.x * yr + psi.y * yi;
        im[c] = psi.y * yr - psi.x * yi;
    }

    const double* row = p.angles + be * p.stride;
    int bad = 0, took1 = 0;
    double c_mine = 1.0, s_mine = 0.0;
    for (int m = 0; m < t.n_steps; ++m) {
        const int within = m % N;
        if (within == 0) {  // lane j of the sample evaluates (cos, sin) of measurement m + j
            const int mm = m + (int)r;
            c_mine = 1.0;
            s_mine = 0.0;
            if (mm < t.n_steps) {
                const StepDev sx = p.steps[mm];
                c_mine = sx.fc;
                s_mine = sx.fs;
                if (sx.angle_idx >= 0) sincos_cw(__ldg(row + sx.angle_idx), s_mine, c_mine);
            }
        }
        const double c = __shfl_sync(0xffffffffu, c_mine, lane_base + within);
        const double s = __shfl_sync(0xffffffffu, s_mine, lane_base + within);
        const StepDev st = p.steps[m];
        const MeasCoef q = meas_coef(st.plane, c, s, t);
        const bool append = (st.flags & MBQC_STEP_APPEND) != 0;
        int outcome = 0;
        switch (st.slot) {
            case 0: dm_reg_stage<W, 0>(re, im, q, r, st.nbr_mask, lane_base, append, outcome, bad); break;
            case 1: if constexpr (W > 1) dm_reg_stage<W, 1>(re, im, q, r, st.nbr_mask, lane_base, append, outcome, bad); break;
            case 2: if constexpr (W > 2) dm_reg_stage<W, 2>(re, im, q, r, st.nbr_mask, lane_base, append, outcome, bad); break;
            case 3: if constexpr (W > 3) dm_reg_stage<W, 3>(re, im, q, r, st.nbr_mask, lane_base, append, outcome, bad); break;
            default: break;
        }
        took1 |= outcome;
        if (live && p.outcomes && r == 0) p.outcomes[b * t.n_steps + m] = (int8_t)outcome;
    }

    // ---- output phase through shared memory: rows -> stage, channel on output qubits, gather ----
    double2* rho = stage + (size_t)ls * N * N;
#pragma unroll
    for (int c = 0; c < N; ++c) rho[r * N + c] = make_double2(re[c], im[c]);
    __syncwarp();
    constexpr uint32_t ngroups = (N * N) >> 2;
    constexpr uint32_t gmask = (N >> 1) - 1;
    if (t.has_noise) {
        const mbqc_noise& nz = t.noise;
        for (int qo = 0; qo < t.n_out; ++qo) {
            const int sl = t.out_slot[qo];
            const uint32_t cbit = 1u << sl, rbit = cbit << W;
            for (uint32_t g = r; g < ngroups; g += N) {
                const uint32_t r0 = (uint32_t)insert_zero(g >> (W - 1), sl);
                const uint32_t c0 = (uint32_t)insert_zero(g & gmask, sl);
                const uint32_t i00 = (r0 << W) | c0;
                const double2 a = rho[i00], bq = rho[i00 | cbit], cq = rho[i00 | rbit], d = rho[i00 | rbit | cbit];
                rho[i00] = make_double2(nz.pop[0] * a.x + nz.pop[1] * d.x, nz.pop[0] * a.y + nz.pop[1] * d.y);
                rho[i00 | rbit | cbit] = make_double2(nz.pop[2] * a.x + nz.pop[3] * d.x, nz.pop[2] * a.y + nz.pop[3] * d.y);
                rho[i00 | cbit] = make_double2(nz.coh_g * bq.x + nz.coh_d * cq.x, nz.coh_g * bq.y + nz.coh_d * cq.y);
                rho[i00 | rbit] = make_double2(nz.coh_g * cq.x + nz.coh_d * bq.x, nz.coh_g * cq.y + nz.coh_d * bq.y);
            }
            __syncwarp();
        }
    }
    const uint32_t no = 1u << t.n_out;
    double tr = 0.0;
    for (uint32_t o = r; o < no; o += N) {
        const uint32_t idx = (uint32_t)output_state_index(t, o);
        tr += rho[(idx << W) | idx].x;
    }
#pragma unroll
    for (int o = N >> 1; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
    if (!live) return;
    if (!(tr > 0.0) || !isfinite(tr)) bad = 1;
    if (p.status && r == 0) p.status[b] = (bad ? MBQC_STATUS_BAD_NORM : 0) | (took1 ? MBQC_STATUS_OUTCOME1 : 0);
    const double sc = 1.0 / tr;
    double2* o = p.out + (b << (2 * t.n_out));
    for (uint32_t e = r; e < no * no; e += N) {
        const uint32_t ri = (uint32_t)output_state_index(t, e >> t.n_out);
        const uint32_t ci = (uint32_t)output_state_index(t, e & (no - 1));
        const double2 v = rho[(ri << W) | ci];
        o[e] = make_double2(v.x * sc, v.y * sc);
    }
}

}  // namespace mbqc
