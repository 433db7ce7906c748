// Batched density-matrix pattern kernel (+ Kraus noise): whole pattern in ONE launch, rho of each
// sample resident in shared memory as a vector over (row bits << w | col bits).
//
// Replaces NumpySimulatorDM.run / measure / measure_ment / reset and their helpers
// (mentpy/simulators/np_simulator_dm.py:151-346, calculator/state_ops.py:16-39,77-119,
// operators/gates.py:62-72,127-143).  Per measurement of slot s, for every 4-group
// (rho00, rho01, rho10, rho11) on (row bit s, col bit s):
//
//     sigma = q00 rho00 + q11 rho11 + q01 rho01 + conj(q01) rho10     tr_s(P E(rho))
//     prob  = sum of Re sigma over the groups on the diagonal          Re tr(rho P)
//     outcome 1 (P -> I - P) iff prob0 < 1e-4                          np_simulator_dm.py:335-338
//     rho'_{ab} = sigma / (2 prob) * sign(r,a) sign(c,b)               |+><+| append + CZ phases
//
// where P = (I + M)/2 is the projector of the measurement plane (ment.py:228-260) and E the
// optional single-qubit channel of the plan folded into the q coefficients.
#pragma once
#include "common.cuh"
#include "dm_params.cuh"

namespace mbqc {

struct MeasCoef {
    double q00, q11, q01r, q01i;
};

// entries of the outcome-0 projector P of plane/angle with the (trace-preserving) channel folded in
__device__ __forceinline__ MeasCoef meas_coef(int plane, double c, double s, const PlanTables& t, double z = 0.0,
                                              bool z_project = false) {
    double p00, p11, p10r, p10i;
    if (plane == MBQC_PLANE_XYZ) {  // axis (c, s, z) given directly (fixed angles)
        p00 = 0.5 * (1.0 + z); p11 = 0.5 * (1.0 - z); p10r = 0.5 * c; p10i = 0.5 * s;
    } else if (plane == MBQC_PLANE_XY) {
        p00 = 0.5; p11 = 0.5; p10r = 0.5 * c; p10i = 0.5 * s;
    } else if (plane == MBQC_PLANE_XZ) {
        p00 = 0.5 * (1.0 + s); p11 = 0.5 * (1.0 - s); p10r = 0.5 * c; p10i = 0.0;
    } else if (plane == MBQC_PLANE_Z) {
        // expectation mode: P0 + P1 = I, the qubit is only traced out; sample mode: P0 = |0><0|
        p00 = 1.0; p11 = z_project ? 0.0 : 1.0; p10r = 0.0; p10i = 0.0;
    } else {
        p00 = 0.5 * (1.0 + s); p11 = 0.5 * (1.0 - s); p10r = 0.0; p10i = 0.5 * c;
    }
    MeasCoef q;
    if (t.has_noise) {
        const mbqc_noise& nz = t.noise;
        q.q00 = p00 * nz.pop[0] + p11 * nz.pop[2];
        q.q11 = p00 * nz.pop[1] + p11 * nz.pop[3];
        // coefficient of rho01: p10 g + p01 d, p01 = conj(p10)
        q.q01r = p10r * (nz.coh_g + nz.coh_d);
        q.q01i = p10i * (nz.coh_g - nz.coh_d);
    } else {
        q.q00 = p00; q.q11 = p11; q.q01r = p10r; q.q01i = p10i;
    }
    return q;
}

__device__ __forceinline__ double2 group_sigma(const MeasCoef& q, double2 r00, double2 r01,
                                               double2 r10, double2 r11) {
    double2 sg;
    // q01 * r01 + conj(q01) * r10
    sg.x = fma(q.q00, r00.x, q.q11 * r11.x);
    sg.y = fma(q.q00, r00.y, q.q11 * r11.y);
    sg.x = fma(q.q01r, r01.x + r10.x, fma(-q.q01i, r01.y - r10.y, sg.x));
    sg.y = fma(q.q01r, r01.y + r10.y, fma(q.q01i, r01.x - r10.x, sg.y));
    return sg;
}

constexpr int kDmMaxGroupsPerThread = 4;

__device__ __forceinline__ double dm_group_sum(double v, int tps_log2, int ls, double* red) {
    if (tps_log2 <= 5) {
        for (int o = (1 << tps_log2) >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const int w0 = (ls << tps_log2) >> 5, nw = 1 << (tps_log2 - 5);
    double tot = 0.0;
    for (int k = 0; k < nw; ++k) tot += red[w0 + k];
    return tot;
}

// barrier over the threads of one sample: a warp-level sync is enough when the group fits a warp
__device__ __forceinline__ void dm_group_barrier(int tps_log2) {
    if (tps_log2 <= 5) __syncwarp();
    else __syncthreads();
}

// TPS = 2^tps_log2 threads per sample (>= 4^(w-1)/kDmMaxGroupsPerThread), SPB samples per CTA.
__global__ void dm_smem_kernel(const __grid_constant__ DmBatchParams p, int tps_log2, int spb) {
    extern __shared__ double2 smem[];
    __shared__ double red[32];
    const PlanTables& t = p.tab;
    const int w = t.window;
    const int tps = 1 << tps_log2;
    const int ls = threadIdx.x >> tps_log2;
    const int tid = threadIdx.x & (tps - 1);
    const int64_t b = (int64_t)blockIdx.x * spb + ls;
    const bool live = b < p.batch;
    const uint32_t dim = 1u << w;
    const uint32_t nelem = 1u << (2 * w);
    // per sample: rho [4^w] followed by a scratch vector psi [2^w]
    double2* rho = smem + (size_t)ls * (nelem + dim);
    double2* psi = rho + nelem;

    if (live) {
        const double2* in = (p.input_mode == MBQC_INPUT_PLUS)
                                ? nullptr
                                : p.inputs + (p.input_mode == MBQC_INPUT_BATCH ? (b << t.n_in) : 0);
        const double a0 = t.plus_amp;
        for (uint32_t i = tid; i < dim; i += tps) {
            double2 v = make_double2(a0, 0.0);
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
        }
    }
    dm_group_barrier(tps_log2);
    if (live)
        for (uint32_t e = tid; e < nelem; e += tps) {
            const double2 x = psi[e >> w], y = psi[e & (dim - 1)];
            rho[e] = make_double2(x.x * y.x + x.y * y.y, x.y * y.x - x.x * y.y);
        }
    dm_group_barrier(tps_log2);

    const double* row = p.angles + (live ? b : 0) * p.stride;
    const uint32_t ngroups = nelem >> 2;
    const uint32_t gmask = (dim >> 1) - 1;  // w-1 bits
    int bad = 0, took1 = 0;
    uint32_t hist = 0;  // outcome record, newest in bit 0 (controlled steps)
    // (cos, sin) of the next `chunk` measurements are evaluated by different lanes of the sample's
    // group (one sincos per measurement per warp instead of one per lane) and broadcast by shuffle
    const int chunk = tps < 32 ? tps : 32;
    const int lane = threadIdx.x & 31;
    const int lane_base = tps < 32 ? (lane & ~(tps - 1)) : 0;
    double c_mine = 1.0, s_mine = 0.0;
    for (int m = 0; m < t.n_steps; ++m) {
        const int within = m % chunk;
        if (within == 0) {
            const int mm = m + (tps < 32 ? tid : lane);
            c_mine = 1.0;
            s_mine = 0.0;
            if (mm < t.n_steps) {
                const StepDev sx = p.steps[mm];
                c_mine = sx.fc;
                s_mine = sx.fs;
                const int ai = sx.angle_idx >= 0 ? sx.angle_idx : sx.alt_angle_idx;  // controlled step: either branch
                if (ai >= 0) sincos_cw(__ldg(row + ai), s_mine, c_mine);
            }
        }
        double c = __shfl_sync(0xffffffffu, c_mine, lane_base + within);
        double s = __shfl_sync(0xffffffffu, s_mine, lane_base + within);
        const StepDev st = p.steps[m];
        int plane = st.plane;
        double fz = st.fz;
        if (st.cond_mask) {  // outcome-controlled measurement (controlled_ment.py:96-113)
            const bool alt = cond_takes_alt(hist, st.cond_mask, st.cond_table);
            plane = alt ? st.alt_plane : st.plane;
            fz = alt ? st.afz : st.fz;
            if ((alt ? st.alt_angle_idx : st.angle_idx) < 0) {
                c = alt ? st.afc : st.fc;
                s = alt ? st.afs : st.fs;
            }
        }
        const int sl = st.slot;
        const uint32_t cbit = 1u << sl, rbit = cbit << w;
        const MeasCoef q = meas_coef(plane, c, s, t, fz);
        double2 sg[kDmMaxGroupsPerThread], sf[kDmMaxGroupsPerThread];
        double tr0 = 0.0, trf = 0.0;
        if (live) {
#pragma unroll
            for (int k = 0; k < kDmMaxGroupsPerThread; ++k) {
                const uint32_t g = tid + (uint32_t)k * tps;
                if (g < ngroups) {
                    const uint32_t r0 = (uint32_t)insert_zero(g >> (w - 1), sl);
                    const uint32_t c0 = (uint32_t)insert_zero(g & gmask, sl);
                    const uint32_t i00 = (r0 << w) | c0;
                    const double2 r00 = rho[i00], r11 = rho[i00 | rbit | cbit];
                    sg[k] = group_sigma(q, r00, rho[i00 | cbit], rho[i00 | rbit], r11);
                    sf[k] = make_double2(r00.x + r11.x, r00.y + r11.y);  // tr_s(rho): P0 + P1 = I
                    if (r0 == c0) {
                        tr0 += sg[k].x;
                        trf += sf[k].x;
                    }
                }
            }
        }
        tr0 = dm_group_sum(tr0, tps_log2, ls, red);
        trf = dm_group_sum(trf, tps_log2, ls, red);
        // outcome 1 iff prob0 < 1e-4 (np_simulator_dm.py:335-338); sigma1 = tr_s(rho) - sigma0
        const int outcome = (tr0 < 1e-4) ? 1 : 0;
        took1 |= outcome;
        hist = (hist << 1) | (uint32_t)outcome;
        const double prob = outcome ? (trf - tr0) : tr0;
        if (outcome) {
#pragma unroll
            for (int k = 0; k < kDmMaxGroupsPerThread; ++k)
                sg[k] = make_double2(sf[k].x - sg[k].x, sf[k].y - sg[k].y);
        }
        if (!(prob > 0.0) || !isfinite(prob)) bad = 1;
        if (live && p.outcomes && tid == 0) p.outcomes[b * t.n_steps + m] = (int8_t)outcome;
        const double sc = 0.5 / prob;
        if (live) {
#pragma unroll
            for (int k = 0; k < kDmMaxGroupsPerThread; ++k) {
                const uint32_t g = tid + (uint32_t)k * tps;
                if (g < ngroups) {
                    const uint32_t r0 = (uint32_t)insert_zero(g >> (w - 1), sl);
                    const uint32_t c0 = (uint32_t)insert_zero(g & gmask, sl);
                    const uint32_t i00 = (r0 << w) | c0;
                    const double2 v = make_double2(sg[k].x * sc, sg[k].y * sc);
                    const double2 nv = make_double2(-v.x, -v.y);
                    const uint32_t pr = parity64(r0 & st.nbr_mask), pc = parity64(c0 & st.nbr_mask);
                    rho[i00] = v;
                    rho[i00 | cbit] = pc ? nv : v;
                    rho[i00 | rbit] = pr ? nv : v;
                    rho[i00 | rbit | cbit] = (pr ^ pc) ? nv : v;
                }
            }
        }
        dm_group_barrier(tps_log2);
    }

    // channel on the output qubits (pennylane_simulator.py:123-136 touches every wire)
    if (t.has_noise) {
        const mbqc_noise& nz = t.noise;
        for (int qo = 0; qo < t.n_out; ++qo) {
            const int sl = t.out_slot[qo];
            const uint32_t cbit = 1u << sl, rbit = cbit << w;
            if (live)
                for (uint32_t g = tid; g < ngroups; g += tps) {
                    const uint32_t r0 = (uint32_t)insert_zero(g >> (w - 1), sl);
                    const uint32_t c0 = (uint32_t)insert_zero(g & gmask, sl);
                    const uint32_t i00 = (r0 << w) | c0;
                    const double2 a = rho[i00], bq = rho[i00 | cbit], cq = rho[i00 | rbit], d = rho[i00 | rbit | cbit];
                    rho[i00] = make_double2(nz.pop[0] * a.x + nz.pop[1] * d.x, nz.pop[0] * a.y + nz.pop[1] * d.y);
                    rho[i00 | rbit | cbit] = make_double2(nz.pop[2] * a.x + nz.pop[3] * d.x, nz.pop[2] * a.y + nz.pop[3] * d.y);
                    rho[i00 | cbit] = make_double2(nz.coh_g * bq.x + nz.coh_d * cq.x, nz.coh_g * bq.y + nz.coh_d * cq.y);
                    rho[i00 | rbit] = make_double2(nz.coh_g * cq.x + nz.coh_d * bq.x, nz.coh_g * cq.y + nz.coh_d * bq.y);
                }
            dm_group_barrier(tps_log2);
        }
    }

    // output gather, renormalised by the trace of the gathered block (dead slots carry |+><+|)
    const uint32_t no = 1u << t.n_out;
    double tr = 0.0;
    if (live)
        for (uint32_t o = tid; o < no; o += tps) {
            const uint32_t idx = (uint32_t)output_state_index(t, o);
            tr += rho[(idx << w) | idx].x;
        }
    tr = dm_group_sum(tr, tps_log2, ls, red);
    if (!live) return;
    if (!(tr > 0.0) || !isfinite(tr)) bad = 1;
    if (p.status && tid == 0)
        p.status[b] = (bad ? MBQC_STATUS_BAD_NORM : 0) | (took1 ? MBQC_STATUS_OUTCOME1 : 0);
    const double sc = 1.0 / tr;
    double2* o = p.out + (b << (2 * t.n_out));
    for (uint32_t e = tid; e < no * no; e += tps) {
        const uint32_t ri = (uint32_t)output_state_index(t, e >> t.n_out);
        const uint32_t ci = (uint32_t)output_state_index(t, e & (no - 1));
        const double2 v = rho[(ri << w) | ci];
        o[e] = make_double2(v.x * sc, v.y * sc);
    }
}

}  // namespace mbqc
