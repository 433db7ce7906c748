// Register-resident batched density-matrix kernel (window w <= 5): one LANE per row of rho.
//
// A sample is handled by 2^w lanes of a warp (2 samples per warp at w = 4); lane r keeps row r of
// rho (2^w complex numbers) in registers.  For the measurement of slot s the column bit is a
// register index (compile-time pairs) and the row bit is a lane bit (one __shfl_xor exchange):
//
//     u_a    = q_{a0} rho_{a0} + q_{a1} rho_{a1}        per lane, a = its row bit     (registers)
//     sigma  = u_0 + u_1                                 exchange with lane ^ (1 << s) (shuffles)
//     2 tr   = sum over lanes of Re sigma[r][r]          2^w-lane shuffle reduction
//     rho'_{ab} = sigma * sign(r,a) sign(c,b)            back into the same registers
//
// rho stays unnormalised between steps (trace carried as a scalar, reset every 2^w steps), so a
// step has no division.  No shared-memory traffic, no barriers and no index arithmetic inside the
// step loop: ~1/4 of the instructions of dm_smem_kernel (which stays for w = 6).  Same arithmetic and quirks
// (np_simulator_dm.py:151-346: outcome 1 iff prob0 < 1e-4, per-step normalisation, NaN status),
// optional channel folded into the projector coefficients, channel on the output qubits and the
// output gather done once at the end through a small shared-memory stage.
#pragma once
#include "dm_batch.cuh"
#include "sample.cuh"

namespace mbqc {

// Value of v[pair(r)] for a lane-varying row r: binary select tree over the column bits != S
// (pair indices enumerate the columns with bit S clear in increasing order).
template <int W, int S>
__device__ __forceinline__ double select_diag(const double (&v)[1 << (W - 1)], uint32_t r) {
    constexpr int NP = 1 << (W - 1);
    double sel[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) sel[i] = v[i];
    int n = NP;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        if (j == S) continue;
        const bool hi = (r >> j) & 1u;
        n >>= 1;
#pragma unroll
        for (int i = 0; i < n; ++i) sel[i] = hi ? sel[2 * i + 1] : sel[2 * i];
    }
    return sel[0];
}

// re[r] for a lane-varying r: binary select tree over all column bits
template <int W>
__device__ __forceinline__ double select_own(const double (&v)[1 << W], uint32_t r) {
    double sel[1 << W];
#pragma unroll
    for (int i = 0; i < (1 << W); ++i) sel[i] = v[i];
    int n = 1 << W;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const bool hi = (r >> j) & 1u;
        n >>= 1;
#pragma unroll
        for (int i = 0; i < n; ++i) sel[i] = hi ? sel[2 * i + 1] : sel[2 * i];
    }
    return sel[0];
}

// One measurement of slot S.  rho is kept UNNORMALISED (trace `trc`, uniform over the sample's
// lanes): the new blocks are +-sigma, so trace' = 2 tr(sigma) and no division is needed per step;
// prob0 = tr(sigma0) / trc decides the outcome exactly as the reference's normalised state does.
// rule: kDmRuleThreshold = the reference's deterministic rule; kDmRuleSample: outcome 0 with
// probability prob0 (u = uniform draw); kDmRuleForced: outcome = (u != 0).  pstep receives the
// probability of the outcome taken.
constexpr int kDmRuleThreshold = 0, kDmRuleSample = 1, kDmRuleForced = 2, kDmRuleTrace = 3;  // Trace: plane Z, expectation mode

template <int W, int S>
__device__ __forceinline__ void dm_reg_stage(double (&re)[1 << W], double (&im)[1 << W], const MeasCoef& q,
                                             uint32_t r, uint32_t colpar, double& trc, int& outcome, int& bad,
                                             int rule = kDmRuleThreshold, double u = 0.0, double* pstep = nullptr) {
    constexpr int N = 1 << W;
    constexpr int NP = N >> 1;
    const uint32_t a = (r >> S) & 1u;
    // u = alpha * rho_{a0} + beta * rho_{a1}; a = 0: (q00, q01), a = 1: (conj(q01), q11)
    const double ar = a ? q.q01r : q.q00, ai = a ? -q.q01i : 0.0;
    const double br = a ? q.q11 : q.q01r, bi = a ? 0.0 : q.q01i;
    double sr[NP], si[NP];
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
        ++p;
    }
    // both lanes of a row pair hold the same sigma, so the all-lane sum is 2 tr(sigma0) = trace'
    double t2 = select_diag<W, S>(sr, r);
#pragma unroll
    for (int o = N >> 1; o > 0; o >>= 1) t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    if (rule == kDmRuleThreshold)
        outcome = (0.5 * t2 < 1e-4 * trc) ? 1 : 0;  // prob0 < 1e-4, np_simulator_dm.py:335-338
    else if (rule == kDmRuleSample)
        outcome = (u * trc < 0.5 * t2) ? 0 : 1;
    else if (rule == kDmRuleForced)
        outcome = (u != 0.0) ? 1 : 0;
    else
        outcome = 0;
    if (__any_sync(0xffffffffu, outcome)) {
        // rare: sigma1 = tr_s(rho) - sigma0, tr_s(rho) = rho_00 + rho_11 (own diagonal block + partner's)
        double fr[NP], fi[NP];
        p = 0;
#pragma unroll
        for (int c0 = 0; c0 < N; ++c0) {
            if (c0 & (1 << S)) continue;
            const int c1 = c0 | (1 << S);
            double xr = a ? re[c1] : re[c0], xi = a ? im[c1] : im[c0];
            xr += __shfl_xor_sync(0xffffffffu, xr, 1 << S);
            xi += __shfl_xor_sync(0xffffffffu, xi, 1 << S);
            fr[p] = xr;
            fi[p] = xi;
            ++p;
        }
        double f2 = select_diag<W, S>(fr, r);
#pragma unroll
        for (int o = N >> 1; o > 0; o >>= 1) f2 += __shfl_xor_sync(0xffffffffu, f2, o);
        if (outcome) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                sr[i] = fr[i] - sr[i];
                si[i] = fi[i] - si[i];
            }
            t2 = f2 - t2;
        }
    }
    if (!(t2 > 0.0) || !isfinite(t2)) bad = 1;
    if (pstep) *pstep = 0.5 * t2 / trc;
    trc = t2;
    // CZ signs with the slot's new |+>: bit c of colpar = parity(c & nbr_mask) (0 for tail steps)
    const uint32_t pr = (a & (colpar >> r)) << 31;
    p = 0;
#pragma unroll
    for (int c0 = 0; c0 < N; ++c0) {
        if (c0 & (1 << S)) continue;
        const int c1 = c0 | (1 << S);
        const uint32_t pc = pr ^ ((colpar << (31 - c0)) & 0x80000000u);
        re[c0] = flip_sign(sr[p], pr);
        im[c0] = flip_sign(si[p], pr);
        re[c1] = flip_sign(sr[p], pc);
        im[c1] = flip_sign(si[p], pc);
        ++p;
    }
}

// SAMPLE = false: the reference's deterministic outcome rule (mbqc_run_batch_dm).  SAMPLE = true:
// Born-rule sampling / forced records with flow corrections (mbqc_run_batch_dm_sampled, see
// sample.cuh); with noise every sample is one branch of the noisy pattern, and the
// probability-weighted mean of the corrected outputs is the PennyLane backend's result.
template <int W, bool SAMPLE>
__global__ void __launch_bounds__(128) dm_reg_kernel(const __grid_constant__ DmBatchParams p,
                                                     const __grid_constant__ SampleParams sp) {
    constexpr int N = 1 << W;           // lanes per sample = columns per lane
    constexpr int SPW = 32 / N;         // samples per warp
    const int SPB = (int)(blockDim.x >> 5) * SPW;  // samples per CTA (1..4 warps, chosen by the launcher)
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
    uint32_t hist = 0, bx = 0, bz = 0;  // SAMPLE: last 32 outcomes, output byproduct bits
    double pb = 1.0;                    // SAMPLE: probability of the outcome record
    double c_mine = 1.0, s_mine = 0.0;
    double trc = psi.x * psi.x + psi.y * psi.y;  // trace of the (unnormalised) rho
#pragma unroll
    for (int o = N >> 1; o > 0; o >>= 1) trc += __shfl_xor_sync(0xffffffffu, trc, o);
    for (int m = 0; m < t.n_steps; ++m) {
        const int within = m % N;
        if (within == 0) {  // lane j of the sample evaluates (cos, sin) of measurement m + j
            if (m) {  // bring the trace back to 1 (it moves by 2 prob per step)
                const double sc = 1.0 / trc;
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    re[c] *= sc;
                    im[c] *= sc;
                }
                trc = 1.0;
            }
            const int mm = m + (int)r;
            c_mine = 1.0;
            s_mine = 0.0;
            if (mm < t.n_steps) {
                const StepDev sx = p.steps[mm];
                c_mine = sx.fc;
                s_mine = sx.fs;
                // a controlled step stages the trainable angle of whichever branch has one
                const int ai = sx.angle_idx >= 0 ? sx.angle_idx : sx.alt_angle_idx;
                if (ai >= 0) sincos_cw(__ldg(row + ai), s_mine, c_mine);
            }
        }
        double c = __shfl_sync(0xffffffffu, c_mine, lane_base + within);
        double s = __shfl_sync(0xffffffffu, s_mine, lane_base + within);
        const StepDev st = p.steps[m];
        int plane = st.plane;
        double fz = st.fz;
        if (st.cond_mask) {  // outcome-controlled measurement (controlled_ment.py:96-113): pick the branch
            const bool alt = cond_takes_alt(hist, st.cond_mask, st.cond_table);
            plane = alt ? st.alt_plane : st.plane;
            fz = alt ? st.afz : st.fz;
            if ((alt ? st.alt_angle_idx : st.angle_idx) < 0) {
                c = alt ? st.afc : st.fc;
                s = alt ? st.afs : st.fs;
            }
        }
        FeedForwardDev ff = {0, 0, 0, 0};
        int rule = kDmRuleThreshold;
        double u = 0.0, pstep = 1.0;
        if constexpr (SAMPLE) {  // theta' = (-1)^a theta + b pi
            ff = sp.ff[m];
            const uint32_t a = __popc(hist & ff.xdep) & 1u, z = __popc(hist & ff.zdep) & 1u;
            c = flip_sign(c, z << 31);
            s = flip_sign(s, (a ^ z) << 31);
            if (sp.outcome_mode == MBQC_OUTCOMES_FORCED) {
                rule = kDmRuleForced;
                u = sp.outcomes[be * t.n_steps + m] ? 1.0 : 0.0;
            } else {
                rule = kDmRuleSample;
                u = philox_uniform(sp.seed, sp.sample_offset + (uint64_t)be, (uint32_t)m);
            }
        }
        const bool zs = p.z_sample != 0;
        const MeasCoef q = meas_coef(plane, c, s, t, fz, zs);
        if (st.plane == MBQC_PLANE_Z && zs) {
            // mode="sample": the reference draws the outcome even under force0 (np_simulator_dm.py:329-333)
            rule = kDmRuleSample;
            u = philox_uniform(p.z_seed, p.z_offset + (uint64_t)be, (uint32_t)m);
        } else if (st.plane == MBQC_PLANE_Z) {
            // expectation mode (np_simulator_dm.py:327-344): record prob1 = tr(P1 E(rho)) / tr(rho)
            // (P1 = |1><1| seen through the channel: weights pop[2], pop[3] on rho00, rho11)
            rule = kDmRuleTrace;
            const bool one = (r >> st.slot) & 1u;
            const double w1 = t.has_noise ? (one ? t.noise.pop[3] : t.noise.pop[2]) : (one ? 1.0 : 0.0);
            double p1 = w1 * select_own<W>(re, r);
#pragma unroll
            for (int o = N >> 1; o > 0; o >>= 1) p1 += __shfl_xor_sync(0xffffffffu, p1, o);
            if (live && r == 0 && p.expect) p.expect[b * t.n_steps + m] = p1 / trc;
        }
        // column-parity table of this step's CZ mask, one bit per lane of the sample (0 for tail steps)
        const bool odd = (st.flags & MBQC_STEP_APPEND) && parity64((uint64_t)r & st.nbr_mask);
        const uint32_t colpar = __ballot_sync(0xffffffffu, odd) >> lane_base;
        int outcome = 0;
        switch (st.slot) {
            case 0: dm_reg_stage<W, 0>(re, im, q, r, colpar, trc, outcome, bad, rule, u, &pstep); break;
            case 1: if constexpr (W > 1) dm_reg_stage<W, 1>(re, im, q, r, colpar, trc, outcome, bad, rule, u, &pstep); break;
            case 2: if constexpr (W > 2) dm_reg_stage<W, 2>(re, im, q, r, colpar, trc, outcome, bad, rule, u, &pstep); break;
            case 3: if constexpr (W > 3) dm_reg_stage<W, 3>(re, im, q, r, colpar, trc, outcome, bad, rule, u, &pstep); break;
            case 4: if constexpr (W > 4) dm_reg_stage<W, 4>(re, im, q, r, colpar, trc, outcome, bad, rule, u, &pstep); break;
            default: break;
        }
        took1 |= outcome;
        if constexpr (SAMPLE) {
            hist = (hist << 1) | (uint32_t)outcome;
            if (outcome) {
                bx ^= ff.outx;
                bz ^= ff.outz;
            }
            pb *= pstep;
            if (live && r == 0 && sp.outcomes && sp.outcome_mode != MBQC_OUTCOMES_FORCED)
                sp.outcomes[b * t.n_steps + m] = (int8_t)outcome;
        } else {
            hist = (hist << 1) | (uint32_t)outcome;
            if (live && p.outcomes && r == 0) p.outcomes[b * t.n_steps + m] = (int8_t)outcome;
        }
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
    uint32_t xm = 0, zm = 0;
    if constexpr (SAMPLE) {
        if (r == 0) {
            if (sp.byproducts) sp.byproducts[b] = bx | (bz << 16);
            if (sp.prob) sp.prob[b] = pb;
        }
        if (sp.correct) {  // rho' = X^x Z^z rho Z^z X^x on the output register
            xm = byproduct_index_mask(bx, t.n_out);
            zm = byproduct_index_mask(bz, t.n_out);
        }
    }
    for (uint32_t e = r; e < no * no; e += N) {
        const uint32_t dr = (e >> t.n_out) ^ xm, dc = (e & (no - 1)) ^ xm;
        const uint32_t ri = (uint32_t)output_state_index(t, dr);
        const uint32_t ci = (uint32_t)output_state_index(t, dc);
        const double2 v = rho[(ri << W) | ci];
        const uint32_t sg = (uint32_t)(__popc((dr ^ dc) & zm) & 1) << 31;
        o[e] = make_double2(flip_sign(v.x * sc, sg), flip_sign(v.y * sc, sg));
    }
}

}  // namespace mbqc
