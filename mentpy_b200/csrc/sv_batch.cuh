// Batched state-vector pattern kernels: the whole pattern (all measurements) in ONE launch.
//
//   sv_reg_kernel<W>   w <= 5: one thread per angle set, the 2^w amplitudes live in registers
//   sv_smem_kernel     6 <= w <= 12: one thread group per angle set, amplitudes in shared memory
//
// Replaces NumpySimulatorSV.run / measure / measure_ment / reset and the helpers they call
// (mentpy/simulators/np_simulator_sv.py:164-358, calculator/state_ops.py:42-74,
// operators/gates.py:62-72,127-143) -- see common.cuh for the per-measurement identity.
#pragma once
#include "common.cuh"

namespace mbqc {

struct SvBatchParams {
    PlanTables tab;
    const StepDev* __restrict__ steps;
    const double* __restrict__ angles;  // [B][stride]
    int64_t stride;
    const double2* __restrict__ inputs;
    int32_t input_mode;
    int64_t batch;
    double2* __restrict__ out;  // [B][2^k]
    int32_t* __restrict__ status;
    int32_t* __restrict__ status_any;  // optional: OR of all status bits (host pipeline)
    // parameter-shift support (grad kernels): unused by the plain run
    const double2* __restrict__ target;
    double shift;
    double* __restrict__ grad;
    double* __restrict__ cost;
};

// ---- one measurement on a register-resident window ---------------------------------------------
template <int W, int S>
__device__ __forceinline__ void reg_stage(double (&re)[1 << W], double (&im)[1 << W], double c,
                                          double s, uint32_t flipmask) {
#pragma unroll
    for (int i = 0; i < (1 << W); ++i) {
        if (i & (1 << S)) continue;
        const int j = i | (1 << S);
        // t = a_i + (c - i s) a_j
        const double tr = fma(c, re[j], fma(s, im[j], re[i]));
        const double ti = fma(c, im[j], fma(-s, re[j], im[i]));
        re[i] = tr;
        im[i] = ti;
        const uint32_t sb = (flipmask << (31 - j)) & 0x80000000u;
        re[j] = flip_sign(tr, sb);
        im[j] = flip_sign(ti, sb);
    }
}

template <int W>
__device__ __forceinline__ void reg_step(double (&re)[1 << W], double (&im)[1 << W], int slot,
                                         double c, double s, uint32_t flipmask) {
    switch (slot) {
        case 0: reg_stage<W, 0>(re, im, c, s, flipmask); break;
        case 1: if constexpr (W > 1) reg_stage<W, 1>(re, im, c, s, flipmask); break;
        case 2: if constexpr (W > 2) reg_stage<W, 2>(re, im, c, s, flipmask); break;
        case 3: if constexpr (W > 3) reg_stage<W, 3>(re, im, c, s, flipmask); break;
        case 4: if constexpr (W > 4) reg_stage<W, 4>(re, im, c, s, flipmask); break;
        default: break;
    }
}

// Angle sources for the evolve loop.  Global: one dependent DRAM load + sincos per step (fallback
// for very long angle vectors).  Staged: the CTA has already turned its [threads x T] angle tile
// into (cos, sin) pairs in shared memory with coalesced loads and independent sincos evaluations,
// so a step costs one LDS; a parameter shift is a rotation by (cos s, sin s), no second sincos.
struct AngleGlobal {
    const double* row;
    int shift_col;
    double shift;
    __device__ __forceinline__ void get(int idx, double& c, double& s) const {
        double th = __ldg(row + idx);
        if (idx == shift_col) th += shift;
        sincos(th, &s, &c);
    }
};
struct AngleStaged {
    const double2* col0;  // shared memory: (cos, sin) of angle column j at col0[j * pitch]
    int pitch;
    int shift_col;
    double cs, ss;  // cos / sin of the shift
    __device__ __forceinline__ void get(int idx, double& c, double& s) const {
        const double2 v = col0[idx * pitch];
        c = v.x;
        s = v.y;
        if (idx == shift_col) {
            c = v.x * cs - v.y * ss;
            s = v.y * cs + v.x * ss;
        }
    }
};

// Evolve one sample through the whole pattern.  Returns the squared norm over the output entries;
// (zr, zi) accumulates the unnormalised reference phase prod (1 + e^{i theta}).
template <int W, class AngleSrc>
__device__ __forceinline__ double sv_reg_evolve(const SvBatchParams& p, const StepDev* __restrict__ steps,
                                                int64_t b, const AngleSrc& ang, double (&re)[1 << W],
                                                double (&im)[1 << W], double& zr, double& zi) {
    constexpr int N = 1 << W;
    const PlanTables& t = p.tab;
    if (p.input_mode == MBQC_INPUT_PLUS) {
        const double a = t.init_scale * exp2(-0.5 * t.n_in);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            re[i] = flip_sign(a, (t.init_sign << (31 - i)) & 0x80000000u);
            im[i] = 0.0;
        }
    } else {
        const double2* in = p.inputs + (p.input_mode == MBQC_INPUT_BATCH ? (b << t.n_in) : 0);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double2 v = __ldg(in + t.init_src[i]);
            const uint32_t sb = (t.init_sign << (31 - i)) & 0x80000000u;
            re[i] = flip_sign(v.x * t.init_scale, sb);
            im[i] = flip_sign(v.y * t.init_scale, sb);
        }
    }
    zr = 1.0;
    zi = 0.0;
    const int M = t.n_steps;
    for (int m = 0; m < M; ++m) {
        const StepDev st = steps[m];
        double c = st.fc, s = st.fs;
        if (st.angle_idx >= 0) ang.get(st.angle_idx, c, s);
        // reference global phase factor (1 + e^{i theta}), normalised at the end
        const double pr = 1.0 + c, pi = s;
        const double nzr = zr * pr - zi * pi;
        zi = zr * pi + zi * pr;
        zr = nzr;
        reg_step<W>(re, im, st.slot, c, s, st.flipmask);
        if ((m & 15) == 15) {  // keep magnitudes bounded on long patterns
            double n2 = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) n2 = fma(re[i], re[i], fma(im[i], im[i], n2));
            const double r = rsqrt(n2);
            const double rz = rsqrt(zr * zr + zi * zi);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                re[i] *= r;
                im[i] *= r;
            }
            zr *= rz;
            zi *= rz;
        }
    }
    double n2 = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (t.out_dst[i] >= 0) n2 = fma(re[i], re[i], fma(im[i], im[i], n2));
    return n2;
}

// Shared-memory staging done by every CTA of the register kernels before the evolve loop.
//   s_steps [n_steps]        the plan's step records (one coalesced copy instead of a dependent
//                            global load per step)
//   s_raw   [samples][T]     the CTA's angle tile, fetched with cp.async so that ALL of its loads
//                            are in flight together (one DRAM latency for the whole pattern)
//   s_cs    [T][pitch]       (cos, sin) per angle, column-major: thread `row` converts its own row
//                            (independent sincos evaluations -> ILP) and later reads
//                            s_cs[col * pitch + row], conflict-free across the warp
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ void stage_steps(const SvBatchParams& p, StepDev* s_steps) {
    const char* src = reinterpret_cast<const char*>(p.steps);
    char* dst = reinterpret_cast<char*>(s_steps);
    for (int i = threadIdx.x; i < p.tab.n_steps * 3; i += blockDim.x) cp_async16(dst + 16 * i, src + 16 * i);
}

// rows [b0, b0+samples) of the angle matrix -> s_raw[samples][T]
__device__ __forceinline__ void stage_raw_angles(const SvBatchParams& p, double* s_raw, int64_t b0, int samples) {
    const int T = p.tab.n_angles;
    if (p.stride == T) {
        const double* src = p.angles + b0 * T;
        const int total = samples * T;
        for (int i = threadIdx.x; i < total; i += blockDim.x) cp_async8(s_raw + i, src + i);
    } else {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int r = wid; r < samples; r += nw)
            for (int j = lane; j < T; j += 32) cp_async8(s_raw + r * T + j, p.angles + (b0 + r) * p.stride + j);
    }
}

// thread `row` turns its own angle row into (cos, sin) pairs, column-major with pitch `pitch`
__device__ __forceinline__ void convert_own_row(const double* s_raw, double2* s_cs, int T, int row, int pitch) {
#pragma unroll 2
    for (int j = 0; j < T; ++j) {
        double sn, cs;
        sincos(s_raw[row * T + j], &sn, &cs);
        s_cs[j * pitch + row] = make_double2(cs, sn);
    }
}

// DM = false: out is [B][2^k] amplitudes.  DM = true: out is [B][2^k][2^k] = |psi><psi|
// (np_simulator_sv.py:292-293, the reference's default output form); the CTA stages its
// normalised amplitudes in shared memory and writes the outer products fully coalesced.
// Dynamic shared memory: [steps | (cos,sin) tile] during the evolve, re-used as the DM stage.
template <int W, bool DM>
__global__ void __launch_bounds__(128) sv_reg_kernel(const __grid_constant__ SvBatchParams p, int staged) {
    constexpr int N = 1 << W;
    extern __shared__ double2 dyn[];
    const int T = p.tab.n_angles;
    StepDev* s_steps = reinterpret_cast<StepDev*>(dyn);
    double2* s_cs = dyn + 3 * p.tab.n_steps;                                   // [T][blockDim]
    double* s_raw = reinterpret_cast<double*>(s_cs + (size_t)T * blockDim.x);  // [blockDim][T]
    const int64_t b0 = (int64_t)blockIdx.x * blockDim.x;
    const int64_t b = b0 + threadIdx.x;
    const bool live = b < p.batch;
    const int k = p.tab.n_out;
    const int samples = (int)min((int64_t)blockDim.x, p.batch - b0);
    stage_steps(p, s_steps);
    if (staged) stage_raw_angles(p, s_raw, b0, samples);
    cp_async_wait_all();
    __syncthreads();
    double re[N], im[N], zr = 1.0, zi = 0.0, n2 = 1.0;
    if (live) {
        if (staged) {
            convert_own_row(s_raw, s_cs, T, threadIdx.x, blockDim.x);  // own row only: no barrier needed
            const AngleStaged ang{s_cs + threadIdx.x, (int)blockDim.x, -1, 1.0, 0.0};
            n2 = sv_reg_evolve<W>(p, s_steps, b, ang, re, im, zr, zi);
        } else {
            const AngleGlobal ang{p.angles + b * p.stride, -1, 0.0};
            n2 = sv_reg_evolve<W>(p, s_steps, b, ang, re, im, zr, zi);
        }
    }
    if constexpr (DM) __syncthreads();  // everyone is done with s_steps / s_cs: re-use as stage
    if (live) {
        const double zn = zr * zr + zi * zi;
        const bool ok = (n2 > 0.0) && (zn > 0.0) && isfinite(n2) && isfinite(zn);
        if (p.status) p.status[b] = ok ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM;
        if (!ok && p.status_any) atomicOr(p.status_any, MBQC_STATUS_BAD_NORM);
        const double r = rsqrt(n2) * rsqrt(zn);
        const double ur = zr * r, ui = zi * r;  // unit phase / norm
        double2* o = DM ? (dyn + ((size_t)threadIdx.x << k)) : (p.out + (b << k));
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int d = p.tab.out_dst[i];
            if (d >= 0) o[d] = make_double2(re[i] * ur - im[i] * ui, re[i] * ui + im[i] * ur);
        }
    }
    if constexpr (DM) {
        __syncthreads();
        const int64_t total = (int64_t)samples << (2 * k);
        double2* o = p.out + (b0 << (2 * k));
        const uint32_t km = (1u << k) - 1u;
        for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
            const double2* sv = dyn + ((e >> (2 * k)) << k);
            const double2 x = sv[(e >> k) & km], y = sv[e & km];
            o[e] = make_double2(x.x * y.x + x.y * y.y, x.y * y.x - x.x * y.y);
        }
    }
}

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
    const uint64_t n = 1ull << w;
    if (live) {
        const double2* in = (p.input_mode == MBQC_INPUT_PLUS)
                                ? nullptr
                                : p.inputs + (p.input_mode == MBQC_INPUT_BATCH ? (b << t.n_in) : 0);
        const double a0 = t.init_scale * exp2(-0.5 * t.n_in);
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
        double c = st.fc, s = st.fs;
        if (st.angle_idx >= 0) sincos(__ldg(row + st.angle_idx), &s, &c);
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
        __syncthreads();
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
