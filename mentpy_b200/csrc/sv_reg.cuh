// Register-resident batched state-vector kernel (window w <= 5): the whole pattern in ONE launch,
// one thread per angle set, the 2^w amplitudes never leave the register file.
//
// Replaces NumpySimulatorSV.run / measure / measure_ment / reset and the helpers they call
// (mentpy/simulators/np_simulator_sv.py:164-358, calculator/state_ops.py:42-74,
// operators/gates.py:62-72,127-143) -- see common.cuh for the per-measurement identity.
//
// What bounds it (ncu, profiles/): HBM traffic is only angles in / amplitudes out (144 B per
// evaluation on grid_cluster(2,6)), the work is ~45 FP64 instructions per measurement, so the
// kernel is FP64-pipe / issue bound.  Design points that follow from that:
//   * all angle loads of a CTA are issued together with cp.async (one DRAM latency per pattern,
//     not one per measurement) and converted to (cos, sin) by independent evaluations;
//   * sincos is a branch-light Cody-Waite + fdlibm-kernel implementation with its constants in
//     the constant bank (FMA operands), ~1/2 the instructions of the library call;
//   * measured slots follow m -> w-1-(m mod w) for every reference schedule (window position 0
//     is always measured and its slot recycled), so the step loop is unrolled by w with
//     compile-time slots: no switch, no register shuffling between cases;
//   * CZ signs come as ready-made sign words (0 / 0x80000000) from shared memory and are applied
//     with one LOP3 per word on the integer pipe, keeping the FP64 pipe for the FMAs.
#pragma once
#include <utility>

#include "sv_batch.cuh"

namespace mbqc {

// device tables for the register kernel, built once per plan
struct RegPlanDev {
    const uint32_t* __restrict__ cols;   // [M] (slot << 16) | column; column >= T: fixed[column-T]
    const uint32_t* __restrict__ signs;  // [M][SP] sign words of the pair partners
    const double2* __restrict__ fixed;   // [n_fixed] (cos, sin) of fixed-angle steps
    int32_t n_fixed;
    int32_t sign_pitch;  // SP = max(2^(w-1), 4)
    int32_t periodic;    // slots follow w-1-(m mod w)
};

struct SvRegParams {
    SvBatchParams base;
    RegPlanDev reg;
};

// ---- angle sources -----------------------------------------------------------------------------
// Staged: the CTA turned its angle tile into (cos, sin) pairs in shared memory (column j of row
// `row` at col0[j * pitch]); a parameter shift is a rotation by (cos s, sin s).  Global: fallback
// for angle vectors too long to stage (one dependent load + sincos per measurement).
struct AngleStaged {
    const double2* col0;
    int pitch;
    int n_angles;
    const double2* fixed;
    int shift_col;
    double cs, ss;
    // FIXED: the plan has fixed-angle steps; SHIFT: one column carries a parameter shift.  Plain
    // runs of all-trainable patterns compile both checks away.
    template <bool FIXED, bool SHIFT>
    __device__ __forceinline__ void get(uint32_t col, double& c, double& s) const {
        if constexpr (FIXED) {
            if ((int)col >= n_angles) {
                const double2 f = fixed[col - n_angles];
                c = f.x;
                s = f.y;
                return;
            }
        }
        const double2 v = col0[col * pitch];
        c = v.x;
        s = v.y;
        if constexpr (SHIFT) {
            if ((int)col == shift_col) {
                c = v.x * cs - v.y * ss;
                s = v.y * cs + v.x * ss;
            }
        }
    }
};
struct AngleGlobal {
    const double* row;
    int n_angles;
    const double2* fixed;
    int shift_col;
    double shift;
    template <bool FIXED, bool SHIFT>
    __device__ __forceinline__ void get(uint32_t col, double& c, double& s) const {
        if ((int)col >= n_angles) {
            const double2 f = fixed[col - n_angles];
            c = f.x;
            s = f.y;
            return;
        }
        double th = __ldg(row + col);
        if ((int)col == shift_col) th += shift;
        sincos_cw(th, s, c);
    }
};

template <class F, int... U>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, U...>) {
    (f(std::integral_constant<int, U>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

// ---- one measurement ---------------------------------------------------------------------------
template <int W, int S>
__device__ __forceinline__ void reg_stage(double (&re)[1 << W], double (&im)[1 << W], double c,
                                          double s, const uint32_t* __restrict__ sg) {
    constexpr int NP = 1 << (W - 1);
    uint32_t w[NP < 4 ? 4 : NP];
#pragma unroll
    for (int q = 0; q < (NP < 4 ? 1 : NP / 4); ++q) {
        const uint4 v = reinterpret_cast<const uint4*>(sg)[q];
        w[4 * q + 0] = v.x;
        w[4 * q + 1] = v.y;
        w[4 * q + 2] = v.z;
        w[4 * q + 3] = v.w;
    }
    int p = 0;
#pragma unroll
    for (int i = 0; i < (1 << W); ++i) {
        if (i & (1 << S)) continue;
        const int j = i | (1 << S);
        // t = a_i + (c - i s) a_j
        const double tr = fma(c, re[j], fma(s, im[j], re[i]));
        const double ti = fma(c, im[j], fma(-s, re[j], im[i]));
        re[i] = tr;
        im[i] = ti;
        re[j] = flip_sign(tr, w[p]);
        im[j] = flip_sign(ti, w[p]);
        ++p;
    }
}

template <int W>
__device__ __forceinline__ void reg_step_any(double (&re)[1 << W], double (&im)[1 << W], int slot,
                                             double c, double s, const uint32_t* sg) {
    switch (slot) {
        case 0: reg_stage<W, 0>(re, im, c, s, sg); break;
        case 1: if constexpr (W > 1) reg_stage<W, 1>(re, im, c, s, sg); break;
        case 2: if constexpr (W > 2) reg_stage<W, 2>(re, im, c, s, sg); break;
        case 3: if constexpr (W > 3) reg_stage<W, 3>(re, im, c, s, sg); break;
        case 4: if constexpr (W > 4) reg_stage<W, 4>(re, im, c, s, sg); break;
        default: break;
    }
}

template <int W>
__device__ __forceinline__ void reg_renorm(double (&re)[1 << W], double (&im)[1 << W], double& zr, double& zi) {
    double n2 = 0.0;
#pragma unroll
    for (int i = 0; i < (1 << W); ++i) n2 = fma(re[i], re[i], fma(im[i], im[i], n2));
    const double r = rsqrt(n2);
    const double rz = rsqrt(zr * zr + zi * zi);
#pragma unroll
    for (int i = 0; i < (1 << W); ++i) {
        re[i] *= r;
        im[i] *= r;
    }
    zr *= rz;
    zi *= rz;
}

// shared-memory views of the plan tables staged by the CTA
struct RegSmem {
    const uint32_t* cols;
    const uint32_t* signs;
    int sign_pitch;
};

// Evolve one sample through the whole pattern.  Returns the squared norm over the output entries;
// (zr, zi) accumulates the unnormalised reference phase prod_j (1 + e^{i theta_j}).
template <int W, bool FIXED, bool SHIFT, class AngleSrc>
__device__ __forceinline__ double sv_reg_evolve(const SvBatchParams& p, const RegSmem& sm, bool periodic,
                                                int64_t b, const AngleSrc& ang, double (&re)[1 << W],
                                                double (&im)[1 << W], double& zr, double& zi) {
    constexpr int N = 1 << W;
    const PlanTables& t = p.tab;
    if (p.input_mode == MBQC_INPUT_PLUS) {
        const double a = t.plus_amp;
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
    auto phase = [&](double c, double s) {  // (zr, zi) *= (1 + c, s)
        const double pr = 1.0 + c;
        const double nzr = fma(zr, pr, -zi * s);
        zi = fma(zr, s, zi * pr);
        zr = nzr;
    };
    if (periodic) {
        for (int m0 = 0; m0 < M; m0 += W) {
            static_for<W>([&](auto uc) {
                constexpr int u = decltype(uc)::value;
                const int m = m0 + u;
                if (m < M) {
                    double c, s;
                    ang.template get<FIXED, SHIFT>(sm.cols[m] & 0xffffu, c, s);
                    phase(c, s);
                    reg_stage<W, W - 1 - u>(re, im, c, s, sm.signs + m * sm.sign_pitch);
                }
            });
            if (((m0 + W) >> 4) != (m0 >> 4)) reg_renorm<W>(re, im, zr, zi);  // long patterns
        }
    } else {
        for (int m = 0; m < M; ++m) {
            const uint32_t cw = sm.cols[m];
            double c, s;
            ang.template get<FIXED, SHIFT>(cw & 0xffffu, c, s);
            phase(c, s);
            reg_step_any<W>(re, im, (int)(cw >> 16), c, s, sm.signs + m * sm.sign_pitch);
            if ((m & 15) == 15) reg_renorm<W>(re, im, zr, zi);
        }
    }
    double n2 = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (t.out_dst[i] >= 0) n2 = fma(re[i], re[i], fma(im[i], im[i], n2));
    return n2;
}

// ---- shared-memory staging ---------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

#ifndef MBQC_REG_MINBLOCKS_W3
#define MBQC_REG_MINBLOCKS_W3 8
#endif
constexpr int kRegThreads = 128;  // CTA size of the register kernels

// dynamic shared memory layout of the register kernels (all offsets 16-byte aligned)
struct RegSmemLayout {
    uint32_t* signs;   // [M][SP]
    uint32_t* cols;    // [M] padded to 4
    double2* fixed;    // [n_fixed]
    double2* trig;     // [64] copy of kTrigTable for the per-lane gather in sincos_tab
    double2* cs;       // [T][pitch] (cos, sin), column-major   (staged only)
};
__host__ __device__ __forceinline__ size_t reg_smem_tables_bytes(int M, int sp, int n_fixed) {
    return (size_t)M * sp * 4 + (size_t)((M + 3) & ~3) * 4 + (size_t)n_fixed * 16 + (size_t)MBQC_TRIG_N * 16;
}
__device__ __forceinline__ RegSmemLayout reg_smem_carve(void* base, int M, int sp, int n_fixed) {
    RegSmemLayout l;
    char* p = reinterpret_cast<char*>(base);
    l.signs = reinterpret_cast<uint32_t*>(p);
    p += (size_t)M * sp * 4;
    l.cols = reinterpret_cast<uint32_t*>(p);
    p += (size_t)((M + 3) & ~3) * 4;
    l.fixed = reinterpret_cast<double2*>(p);
    p += (size_t)n_fixed * 16;
    l.trig = reinterpret_cast<double2*>(p);
    p += (size_t)MBQC_TRIG_N * 16;
    l.cs = reinterpret_cast<double2*>(p);
    return l;
}

// plan tables -> shared memory (a few hundred bytes; one 16-byte cp.async per thread or less)
__device__ __forceinline__ void stage_reg_tables(const SvRegParams& p, const RegSmemLayout& l) {
    const int M = p.base.tab.n_steps;
    const int nsig = (M * p.reg.sign_pitch) >> 2;
    const int ncol = (M + 3) >> 2;
    for (int i = threadIdx.x; i < nsig; i += kRegThreads)
        cp_async16(reinterpret_cast<uint4*>(l.signs) + i, reinterpret_cast<const uint4*>(p.reg.signs) + i);
    for (int i = threadIdx.x; i < ncol; i += kRegThreads)
        cp_async16(reinterpret_cast<uint4*>(l.cols) + i, reinterpret_cast<const uint4*>(p.reg.cols) + i);
    for (int i = threadIdx.x; i < p.reg.n_fixed; i += kRegThreads) cp_async16(l.fixed + i, p.reg.fixed + i);
    if (threadIdx.x < MBQC_TRIG_N) cp_async16(l.trig + threadIdx.x, kTrigTable + threadIdx.x);
}

// Thread `row` fetches its own angle row straight into the low halves of its (cos, sin) slots:
// T independent 8-byte cp.async per thread, all in flight together (one DRAM latency for the
// whole pattern); a warp covers 32 consecutive rows = one contiguous span of the angle matrix,
// so every fetched sector is fully used.
__device__ __forceinline__ void fetch_own_row(const double* __restrict__ grow, double2* cs_col0, int T, int pitch) {
#pragma unroll 1
    for (int j = 0; j < T; ++j) cp_async8(&cs_col0[j * pitch].x, grow + j);
}

// ... and converts them in place to (cos, sin): independent evaluations, no barrier needed since
// every slot is written and read by the same thread.
__device__ __forceinline__ void convert_own_row(double2* cs_col0, int T, int pitch, const double2* trig) {
#pragma unroll 2
    for (int j = 0; j < T; ++j) {
        double sn, c;
        sincos_tab(cs_col0[j * pitch].x, sn, c, trig);
        cs_col0[j * pitch] = make_double2(c, sn);
    }
}

// ---- the kernel --------------------------------------------------------------------------------
// DM = false: out is [B][2^k] amplitudes.  DM = true: out is [B][2^k][2^k] = |psi><psi|
// (np_simulator_sv.py:292-293, the reference's default output form); the CTA stages its
// normalised amplitudes in shared memory and writes the outer products fully coalesced.
// `staged`: bit 0 = angle tile staged in shared memory, bit 1 = CTA-coalesced output stage.
template <int W, bool DM>
__global__ void __launch_bounds__(128, (W <= 3 ? MBQC_REG_MINBLOCKS_W3 : (W == 4 ? 5 : 3))) sv_reg_kernel(const __grid_constant__ SvRegParams pp, int staged) {
    constexpr int N = 1 << W;
    extern __shared__ double2 dyn[];
    const SvBatchParams& p = pp.base;
    const int T = p.tab.n_angles, M = p.tab.n_steps;
    const RegSmemLayout l = reg_smem_carve(dyn, M, pp.reg.sign_pitch, pp.reg.n_fixed);
    const int64_t b0 = (int64_t)blockIdx.x * kRegThreads;
    const int64_t b = b0 + threadIdx.x;
    const bool live = b < p.batch;
    const int k = p.tab.n_out;
    const int samples = (int)min((int64_t)kRegThreads, p.batch - b0);
    stage_reg_tables(pp, l);
    if ((staged & 1) && live) fetch_own_row(p.angles + b * p.stride, l.cs + threadIdx.x, T, kRegThreads);
    cp_async_wait_all();
    __syncthreads();
    const RegSmem sm{l.cols, l.signs, pp.reg.sign_pitch};
    double re[N], im[N], zr = 1.0, zi = 0.0, n2 = 1.0;
    if (live) {
        if (staged & 1) {
            convert_own_row(l.cs + threadIdx.x, T, kRegThreads, l.trig);
            const AngleStaged ang{l.cs + threadIdx.x, kRegThreads, T, l.fixed, -1, 1.0, 0.0};
            if (pp.reg.n_fixed == 0)
                n2 = sv_reg_evolve<W, false, false>(p, sm, pp.reg.periodic != 0, b, ang, re, im, zr, zi);
            else
                n2 = sv_reg_evolve<W, true, false>(p, sm, pp.reg.periodic != 0, b, ang, re, im, zr, zi);
        } else {
            const AngleGlobal ang{p.angles + b * p.stride, T, l.fixed, -1, 0.0};
            n2 = sv_reg_evolve<W, true, false>(p, sm, pp.reg.periodic != 0, b, ang, re, im, zr, zi);
        }
    }
    // Output.  Device-resident callers get direct 16-byte stores from registers.  With
    // `stage_out` (DM form, or an output buffer in page-locked HOST memory) the CTA first
    // collects its normalised amplitudes in shared memory and then writes its whole block as one
    // contiguous span, 512 B per warp instruction: that is what makes stores over PCIe efficient
    // (156 vs 261 us per 65,536-sample step in scripts/zerocopy_probe2.cu).
    const bool stage_out = DM || (staged & 2);
    if (stage_out) __syncthreads();  // everyone is done with the staged tables: re-use the buffer
    if (live) {
        const double zn = zr * zr + zi * zi;
        const bool ok = (n2 > 0.0) && (zn > 0.0) && isfinite(n2) && isfinite(zn);
        if (p.status) p.status[b] = ok ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM;
        if (!ok && p.status_any) atomicOr(p.status_any, MBQC_STATUS_BAD_NORM);
        const double r = rsqrt(n2 * zn);
        const double ur = zr * r, ui = zi * r;  // unit phase / norm
        double2* o = stage_out ? (dyn + ((size_t)threadIdx.x << k)) : (p.out + (b << k));
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int d = p.tab.out_dst[i];
            if (d >= 0) o[d] = make_double2(re[i] * ur - im[i] * ui, re[i] * ui + im[i] * ur);
        }
    }
    if (!stage_out) return;
    __syncthreads();
    if constexpr (DM) {
        const int64_t total = (int64_t)samples << (2 * k);
        double2* o = p.out + (b0 << (2 * k));
        const uint32_t km = (1u << k) - 1u;
        for (int64_t e = threadIdx.x; e < total; e += kRegThreads) {
            const double2* sv = dyn + ((e >> (2 * k)) << k);
            const double2 x = sv[(e >> k) & km], y = sv[e & km];
            o[e] = make_double2(x.x * y.x + x.y * y.y, x.y * y.x - x.x * y.y);
        }
    } else {
        const int total = samples << k;
        double2* o = p.out + (b0 << k);
        for (int e = threadIdx.x; e < total; e += kRegThreads) o[e] = dyn[e];
    }
}

}  // namespace mbqc
