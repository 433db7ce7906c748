// Batched state-vector pattern kernels: the whole pattern (all measurements) in ONE launch.
//
//   sv_reg_kernel<W>   (sv_reg.cuh) w <= 5: one thread per angle set, amplitudes in registers
//   sv_smem_kernel     (this file) 6 <= w <= 12: one thread group per angle set, amplitudes in
//                      shared memory
//
// Replaces NumpySimulatorSV.run / measure / measure_ment / reset and the helpers they call
// (mentpy/simulators/np_simulator_sv.py:164-358, calculator/state_ops.py:42-74,
// operators/gates.py:62-72,127-143) -- see common.cuh for the per-measurement identity.
#pragma once
#include "common.cuh"
#include "sv_params.cuh"

namespace mbqc {

// ---- shared-memory variant for 6 <= w <= 12 -----------------------------------------------------
// A group of TPS = 2^tps_log2 threads cooperates on one sample; SPB samples per CTA.

// sum over the threads of one sample group (all threads of the CTA must call this)
__device__ __forceinline__ double group_sum(double v, int tps_log2, int ls, double* red) {
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
__device__ __forceinline__ void group_barrier(int tps_log2) {
    if (tps_log2 <= 5) __syncwarp();
    else __syncthreads();
}

// (cos, sin) of the next 2^tps_log2 measurements, one per thread of the sample's group, into the
// group's shared-memory strip `cs` (every thread evaluating every angle was ~80 % of the kernel)
__device__ __forceinline__ void stage_group_angles(const SvBatchParams& p, const double* row, int m, int tid, double2* cs) {
    const int mm = m + tid;
    if (mm < p.tab.n_steps) {
        const StepDev sx = p.steps[mm];
        double c = sx.fc, s = sx.fs;
        if (sx.angle_idx >= 0) sincos_cw(__ldg(row + sx.angle_idx), s, c);
        cs[tid] = make_double2(c, s);
    }
}

// Shared-memory position of amplitude i: the low three index bits (the 16-byte bank group) are XORed
// with every higher 3-bit field, so that the 8 amplitudes of a group AND the groups of neighbouring
// threads land in different bank groups whichever three slots a pass works on (linear over GF(2):
// swz(a ^ b) = swz(a) ^ swz(b); windows <= 12).
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 3) & 7u) ^ ((i >> 6) & 7u) ^ ((i >> 9) & 7u); }

__global__ void sv_smem_kernel(const __grid_constant__ SvBatchParams p, int tps_log2, int spb, int dm_out) {
    extern __shared__ double2 smem[];
    __shared__ double red[32];
    const PlanTables& t = p.tab;
    const int w = t.window;
    const int tps = 1 << tps_log2;
    const int ls = threadIdx.x >> tps_log2;
    const int tid = threadIdx.x & (tps - 1);
    const int64_t b = (int64_t)blockIdx.x * spb + ls;
    const bool live = b < p.batch;
    double2* psi = smem + ((size_t)ls << w);
    double2* cs = smem + ((size_t)spb << w) + ((size_t)ls << tps_log2);  // staged (cos, sin) of 2^tps_log2 steps
    const uint64_t n = 1ull << w;
    if (live) {
        const double2* in = (p.input_mode == MBQC_INPUT_PLUS)
                                ? nullptr
                                : p.inputs + (p.input_mode == MBQC_INPUT_BATCH ? (b << t.n_in) : 0);
        const double a0 = t.plus_amp;
        for (uint64_t i = tid; i < n; i += tps) {
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
            psi[swz((uint32_t)i)] = v;
        }
    }
    __syncthreads();
    double zr = 1.0, zi = 0.0;
    const double* row = p.angles + (live ? b : 0) * p.stride;
    const uint64_t half = n >> 1;
    // Three consecutive measurements per pass over the state where their slots differ (every
    // reference schedule): a thread loads the 8 amplitudes spanned by the three slot bits, applies
    // the three steps in registers and stores them back -- a third of the shared-memory traffic and
    // of the barriers of one pass per step.  Other steps (tails, repeated slots) go one at a time.
    int staged = -(1 << 30);  // first step of the staged (cos, sin) strip
    int m = 0;
    while (m < t.n_steps) {
        const StepDev s0 = p.steps[m];
        int k = 1;
        StepDev s1 = s0, s2 = s0;
        if (m + 3 <= t.n_steps && w >= 3) {
            s1 = p.steps[m + 1];
            s2 = p.steps[m + 2];
            if (s0.slot != s1.slot && s0.slot != s2.slot && s1.slot != s2.slot) k = 3;
        }
        if (m + k > staged + tps) {
            group_barrier(tps_log2);  // the strip is still being read by slower threads of the group
            stage_group_angles(p, row, m, tid, cs);
            staged = m;
            group_barrier(tps_log2);
        }
        if (k == 3) {
            double cc[3], ss[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const double2 v = cs[m + q - staged];
                cc[q] = v.x;
                ss[q] = v.y;
                const double pr = 1.0 + v.x, pi = v.y;
                const double nzr = zr * pr - zi * pi;
                zi = zr * pi + zi * pr;
                zr = nzr;
            }
            const int sl[3] = {s0.slot, s1.slot, s2.slot};
            const uint32_t mk[3] = {(uint32_t)s0.nbr_mask, (uint32_t)s1.nbr_mask, (uint32_t)s2.nbr_mask};
            const int lo = min(sl[0], min(sl[1], sl[2])), hi = max(sl[0], max(sl[1], sl[2]));
            const int mid = sl[0] + sl[1] + sl[2] - lo - hi;
            // per pass: position offset of every member of a group and its share of the three CZ parities
            // (index arithmetic in 32 bits: windows <= 12; both are linear in the index bits)
            uint32_t eoff[8], epar = 0;  // epar bit (q * 8 + e): parity(member bits of e & mask of step q)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t eb = ((uint32_t)(e & 1) << sl[0]) | ((uint32_t)((e >> 1) & 1) << sl[1]) | ((uint32_t)((e >> 2) & 1) << sl[2]);
                eoff[e] = swz(eb);
#pragma unroll
                for (int q = 0; q < 3; ++q) epar |= (uint32_t)(__popc(eb & mk[q]) & 1) << (q * 8 + e);
            }
            if (live) {
                for (uint32_t g = tid; g < (uint32_t)(n >> 3); g += tps) {
                    const uint32_t base = (uint32_t)insert_zero(insert_zero(insert_zero(g, lo), mid), hi);
                    const uint32_t sb = swz(base);
                    uint32_t bpar = 0;  // bit q: parity(base & mask of step q)
#pragma unroll
                    for (int q = 0; q < 3; ++q) bpar |= (uint32_t)(__popc(base & mk[q]) & 1) << q;
                    double2 v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = psi[sb ^ eoff[e]];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            if (e & (1 << q)) continue;
                            const int f = e | (1 << q);
                            double2 tt;
                            tt.x = fma(cc[q], v[f].x, fma(ss[q], v[f].y, v[e].x));
                            tt.y = fma(cc[q], v[f].y, fma(-ss[q], v[f].x, v[e].y));
                            v[e] = tt;
                            const uint32_t sg = (((bpar >> q) ^ (epar >> (q * 8 + e))) & 1u) << 31;
                            v[f] = make_double2(flip_sign(tt.x, sg), flip_sign(tt.y, sg));
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) psi[sb ^ eoff[e]] = v[e];
                }
            }
        } else {
            const double2 csv = cs[m - staged];
            const double c = csv.x, s = csv.y;
            const double pr = 1.0 + c, pi = s;
            const double nzr = zr * pr - zi * pi;
            zi = zr * pi + zi * pr;
            zr = nzr;
            const uint64_t bit = 1ull << s0.slot;
            if (live) {
                for (uint64_t g = tid; g < half; g += tps) {
                    const uint64_t i0 = insert_zero(g, s0.slot);
                    const uint32_t q0 = swz((uint32_t)i0), q1 = swz((uint32_t)(i0 | bit));
                    const double2 a = psi[q0], bb = psi[q1];
                    double2 tt;
                    tt.x = fma(c, bb.x, fma(s, bb.y, a.x));
                    tt.y = fma(c, bb.y, fma(-s, bb.x, a.y));
                    psi[q0] = tt;
                    if (parity64(i0 & s0.nbr_mask)) {
                        tt.x = -tt.x;
                        tt.y = -tt.y;
                    }
                    psi[q1] = tt;
                }
            }
        }
        group_barrier(tps_log2);
        const bool renorm = ((m + k) >> 3) != (m >> 3);  // a step with (m & 7) == 7 lies in this pass
        m += k;
        if (renorm) {  // keep magnitudes bounded on long patterns
            double n2 = 0.0;
            if (live)
                for (uint64_t i = tid; i < n; i += tps) {
                    const double2 v = psi[i];
                    n2 = fma(v.x, v.x, fma(v.y, v.y, n2));
                }
            n2 = group_sum(n2, tps_log2, ls, red);
            const double r = rsqrt(n2);
            if (live)
                for (uint64_t i = tid; i < n; i += tps) {
                    psi[i].x *= r;
                    psi[i].y *= r;
                }
            const double rz = rsqrt(zr * zr + zi * zi);
            zr *= rz;
            zi *= rz;
            __syncthreads();
        }
    }
    // output gather + norm over the output entries
    const uint32_t no = 1u << t.n_out;
    double n2 = 0.0;
    if (live)
        for (uint32_t o = tid; o < no; o += tps) {
            const double2 v = psi[swz((uint32_t)output_state_index(t, o))];
            n2 = fma(v.x, v.x, fma(v.y, v.y, n2));
        }
    n2 = group_sum(n2, tps_log2, ls, red);
    if (!live) return;
    const double zn = zr * zr + zi * zi;
    const bool ok = (n2 > 0.0) && (zn > 0.0) && isfinite(n2) && isfinite(zn);
    if (p.status && tid == 0) p.status[b] = ok ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM;
    if (!ok && p.status_any && tid == 0) atomicOr(p.status_any, MBQC_STATUS_BAD_NORM);
    const double r = rsqrt(n2) * rsqrt(zn);
    const double ur = zr * r, ui = zi * r;
    if (!dm_out) {
        double2* o = p.out + (b << t.n_out);
        for (uint32_t k = tid; k < no; k += tps) {
            const double2 v = psi[swz((uint32_t)output_state_index(t, k))];
            o[k] = make_double2(v.x * ur - v.y * ui, v.x * ui + v.y * ur);
        }
    } else {  // |psi><psi|: the global phase cancels
        const double r2 = r * r * zn;
        double2* o = p.out + (b << (2 * t.n_out));
        for (uint32_t e = tid; e < no * no; e += tps) {
            const double2 x = psi[swz((uint32_t)output_state_index(t, e >> t.n_out))];
            const double2 y = psi[swz((uint32_t)output_state_index(t, e & (no - 1)))];
            o[e] = make_double2((x.x * y.x + x.y * y.y) * r2, (x.y * y.x - x.x * y.y) * r2);
        }
    }
}

}  // namespace mbqc
