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
            psi[i] = v;
        }
    }
    __syncthreads();
    double zr = 1.0, zi = 0.0;
    const double* row = p.angles + (live ? b : 0) * p.stride;
    const uint64_t half = n >> 1;
    for (int m = 0; m < t.n_steps; ++m) {
        const StepDev st = p.steps[m];
        if ((m & (tps - 1)) == 0) {
            stage_group_angles(p, row, m, tid, cs);
            group_barrier(tps_log2);
        }
        const double2 csv = cs[m & (tps - 1)];
        const double c = csv.x, s = csv.y;
        const double pr = 1.0 + c, pi = s;
        const double nzr = zr * pr - zi * pi;
        zi = zr * pi + zi * pr;
        zr = nzr;
        const uint64_t bit = 1ull << st.slot;
        if (live) {
            for (uint64_t g = tid; g < half; g += tps) {
                const uint64_t i0 = insert_zero(g, st.slot);
                const double2 a = psi[i0], bb = psi[i0 | bit];
                double2 tt;
                tt.x = fma(c, bb.x, fma(s, bb.y, a.x));
                tt.y = fma(c, bb.y, fma(-s, bb.x, a.y));
                psi[i0] = tt;
                if (parity64(i0 & st.nbr_mask)) {
                    tt.x = -tt.x;
                    tt.y = -tt.y;
                }
                psi[i0 | bit] = tt;
            }
        }
        group_barrier(tps_log2);
        if ((m & 7) == 7) {  // keep magnitudes bounded on long patterns
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
            const double2 v = psi[output_state_index(t, o)];
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
            const double2 v = psi[output_state_index(t, k)];
            o[k] = make_double2(v.x * ur - v.y * ui, v.x * ui + v.y * ur);
        }
    } else {  // |psi><psi|: the global phase cancels
        const double r2 = r * r * zn;
        double2* o = p.out + (b << (2 * t.n_out));
        for (uint32_t e = tid; e < no * no; e += tps) {
            const double2 x = psi[output_state_index(t, e >> t.n_out)];
            const double2 y = psi[output_state_index(t, e & (no - 1))];
            o[e] = make_double2((x.x * y.x + x.y * y.y) * r2, (x.y * y.x - x.x * y.y) * r2);
        }
    }
}

}  // namespace mbqc
