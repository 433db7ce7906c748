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

// Evolve SPT samples of one thread through the whole pattern, interleaved step by step so that
// their (independent) FP64 dependency chains overlap and the per-step table reads are shared.
// n2[s] receives the squared norm over the output entries; (zr, zi)[s] accumulates the
// unnormalised reference phase prod_j (1 + e^{i theta_j}).
template <int W, int SPT, bool FIXED, bool SHIFT, class AngleSrc>
__device__ __forceinline__ void sv_reg_evolve_multi(const SvBatchParams& p, const RegSmem& sm, bool periodic,
                                                    const int64_t (&b)[SPT], const AngleSrc (&ang)[SPT],
                                                    double (&re)[SPT][1 << W], double (&im)[SPT][1 << W],
                                                    double (&zr)[SPT], double (&zi)[SPT], double (&n2)[SPT]) {
    constexpr int N = 1 << W;
    const PlanTables& t = p.tab;
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
        if (p.input_mode == MBQC_INPUT_PLUS) {
            const double a = t.plus_amp;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                re[q][i] = flip_sign(a, (t.init_sign << (31 - i)) & 0x80000000u);
                im[q][i] = 0.0;
            }
        } else {
            const double2* in = p.inputs + (p.input_mode == MBQC_INPUT_BATCH ? (b[q] << t.n_in) : 0);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double2 v = __ldg(in + t.init_src[i]);
                const uint32_t sb = (t.init_sign << (31 - i)) & 0x80000000u;
                re[q][i] = flip_sign(v.x * t.init_scale, sb);
                im[q][i] = flip_sign(v.y * t.init_scale, sb);
            }
        }
        zr[q] = 1.0;
        zi[q] = 0.0;
    }
    const int M = t.n_steps;
    auto step_all = [&](auto slot_tag, int m, uint32_t cw) {
        const uint32_t* sg = sm.signs + m * sm.sign_pitch;
#pragma unroll
        for (int q = 0; q < SPT; ++q) {
            double c, s;
            ang[q].template get<FIXED, SHIFT>(cw & 0xffffu, c, s);
            const double pr = 1.0 + c;  // (zr, zi) *= (1 + c, s)
            const double nzr = fma(zr[q], pr, -zi[q] * s);
            zi[q] = fma(zr[q], s, zi[q] * pr);
            zr[q] = nzr;
            constexpr int S = decltype(slot_tag)::value;
            if constexpr (S >= 0) reg_stage<W, S>(re[q], im[q], c, s, sg);
            else reg_step_any<W>(re[q], im[q], (int)(cw >> 16), c, s, sg);
        }
    };
    if (periodic) {
        for (int m0 = 0; m0 < M; m0 += W) {
            static_for<W>([&](auto uc) {
                constexpr int u = decltype(uc)::value;
                const int m = m0 + u;
                if (m < M) step_all(std::integral_constant<int, W - 1 - u>{}, m, sm.cols[m]);
            });
            if (((m0 + W) >> 4) != (m0 >> 4)) {  // long patterns: keep magnitudes bounded
#pragma unroll
                for (int q = 0; q < SPT; ++q) reg_renorm<W>(re[q], im[q], zr[q], zi[q]);
            }
        }
    } else {
        for (int m = 0; m < M; ++m) {
            step_all(std::integral_constant<int, -1>{}, m, sm.cols[m]);
            if ((m & 15) == 15) {
#pragma unroll
                for (int q = 0; q < SPT; ++q) reg_renorm<W>(re[q], im[q], zr[q], zi[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (t.out_dst[i] >= 0) acc = fma(re[q][i], re[q][i], fma(im[q][i], im[q][i], acc));
        n2[q] = acc;
    }
}

// single-sample form (gradient kernel)
template <int W, bool FIXED, bool SHIFT, class AngleSrc>
__device__ __forceinline__ double sv_reg_evolve(const SvBatchParams& p, const RegSmem& sm, bool periodic,
                                                int64_t b, const AngleSrc& ang, double (&re)[1 << W],
                                                double (&im)[1 << W], double& zr, double& zi) {
    const int64_t bb[1] = {b};
    const AngleSrc aa[1] = {ang};
    double (&re1)[1][1 << W] = reinterpret_cast<double (&)[1][1 << W]>(re);
    double (&im1)[1][1 << W] = reinterpret_cast<double (&)[1][1 << W]>(im);
    double z1[1], z2[1], n2[1];
    sv_reg_evolve_multi<W, 1, FIXED, SHIFT>(p, sm, periodic, bb, aa, re1, im1, z1, z2, n2);
    zr = z1[0];
    zi = z2[0];
    return n2[0];
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

#ifndef MBQC_REG_SPT_SMALL
#define MBQC_REG_SPT_SMALL 1  // samples per thread for windows <= 3 (2 measured slower on B200: 71.9 vs 62.7 us per 2^20 on C2)
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
#ifndef MBQC_SINCOS_TAB
#define MBQC_SINCOS_TAB 0
#endif
#ifndef MBQC_CONVERT_UNROLL
#define MBQC_CONVERT_UNROLL 2
#endif
#define MBQC_DO_PRAGMA(x) _Pragma(#x)
#define MBQC_UNROLL(n) MBQC_DO_PRAGMA(unroll n)
__device__ __forceinline__ void convert_own_row(double2* cs_col0, int T, int pitch, const double2* trig) {
    MBQC_UNROLL(MBQC_CONVERT_UNROLL)
    for (int j = 0; j < T; ++j) {
        double sn, c;
#if MBQC_SINCOS_TAB
        sincos_tab(cs_col0[j * pitch].x, sn, c, trig);
#else
        sincos_cw(cs_col0[j * pitch].x, sn, c);
#endif
        cs_col0[j * pitch] = make_double2(c, sn);
    }
}

// ---- the kernel --------------------------------------------------------------------------------
// DM = false: out is [B][2^k] amplitudes.  DM = true: out is [B][2^k][2^k] = |psi><psi|
// (np_simulator_sv.py:292-293, the reference's default output form).
// SPT = samples per thread: the CTA covers 128 * SPT samples, thread t owns samples
// b0 + t + q * 128 (so the (cos, sin) column reads stay conflict-free).  SPT = 2 was measured
// SLOWER than 1 on B200 for w = 3 (fewer resident warps outweigh the extra ILP), so 1 is the default;
// likewise the table-driven sincos (MBQC_SINCOS_TAB=1) loses to the polynomial one (profiles/README.md).
// `staged`: bit 0 = angle tile staged in shared memory, bit 1 = CTA-coalesced output stage.
template <int W>
struct RegKernelTraits {
    static constexpr int kSPT = (W <= 3) ? MBQC_REG_SPT_SMALL : 1;
    static constexpr int kMinBlocks = (W <= 3) ? (kSPT == 1 ? 8 : 4) : (W == 4 ? 5 : 3);
};

template <int W, bool DM>
__global__ void __launch_bounds__(128, RegKernelTraits<W>::kMinBlocks)
sv_reg_kernel(const __grid_constant__ SvRegParams pp, int staged) {
    constexpr int N = 1 << W;
    constexpr int SPT = RegKernelTraits<W>::kSPT;
    constexpr int kPitch = kRegThreads * SPT;
    extern __shared__ double2 dyn[];
    const SvBatchParams& p = pp.base;
    const int T = p.tab.n_angles, M = p.tab.n_steps;
    const RegSmemLayout l = reg_smem_carve(dyn, M, pp.reg.sign_pitch, pp.reg.n_fixed);
    const int64_t b0 = (int64_t)blockIdx.x * kPitch;
    const int k = p.tab.n_out;
    const int samples = (int)min((int64_t)kPitch, p.batch - b0);
    int64_t b[SPT];
    bool live[SPT];
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
        b[q] = b0 + threadIdx.x + q * kRegThreads;
        live[q] = b[q] < p.batch;
    }
    stage_reg_tables(pp, l);
    if (staged & 1) {
#pragma unroll
        for (int q = 0; q < SPT; ++q)
            if (live[q]) fetch_own_row(p.angles + b[q] * p.stride, l.cs + threadIdx.x + q * kRegThreads, T, kPitch);
    }
    cp_async_wait_all();
    __syncthreads();
    const RegSmem sm{l.cols, l.signs, pp.reg.sign_pitch};
    double re[SPT][N], im[SPT][N], zr[SPT], zi[SPT], n2[SPT];
    // a thread with a dead second sample simply recomputes its first one (no divergence, results unused)
    int64_t be[SPT];
#pragma unroll
    for (int q = 0; q < SPT; ++q) be[q] = live[q] ? b[q] : b[0];
    if (live[0]) {
        if (staged & 1) {
            AngleStaged ang[SPT];
#pragma unroll
            for (int q = 0; q < SPT; ++q) {
                double2* col0 = l.cs + threadIdx.x + (live[q] ? q : 0) * kRegThreads;
                if (live[q]) convert_own_row(col0, T, kPitch, l.trig);
                ang[q] = AngleStaged{col0, kPitch, T, l.fixed, -1, 1.0, 0.0};
            }
            if (pp.reg.n_fixed == 0)
                sv_reg_evolve_multi<W, SPT, false, false>(p, sm, pp.reg.periodic != 0, be, ang, re, im, zr, zi, n2);
            else
                sv_reg_evolve_multi<W, SPT, true, false>(p, sm, pp.reg.periodic != 0, be, ang, re, im, zr, zi, n2);
        } else {
            AngleGlobal ang[SPT];
#pragma unroll
            for (int q = 0; q < SPT; ++q) ang[q] = AngleGlobal{p.angles + be[q] * p.stride, T, l.fixed, -1, 0.0};
            sv_reg_evolve_multi<W, SPT, true, false>(p, sm, pp.reg.periodic != 0, be, ang, re, im, zr, zi, n2);
        }
    }
    // Output.  Device-resident callers get direct 16-byte stores from registers.  With
    // `stage_out` (DM form, or an output buffer in page-locked HOST memory) the CTA first
    // collects its normalised amplitudes in shared memory and then writes its whole block as one
    // contiguous span, 512 B per warp instruction: that is what makes stores over PCIe efficient
    // (156 vs 261 us per 65,536-sample step in scripts/zerocopy_probe2.cu).
    const bool stage_out = DM || (staged & 2);
    if (stage_out) __syncthreads();  // everyone is done with the staged tables: re-use the buffer
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
        if (!live[q]) continue;
        const double zn = zr[q] * zr[q] + zi[q] * zi[q];
        const bool ok = (n2[q] > 0.0) && (zn > 0.0) && isfinite(n2[q]) && isfinite(zn);
        if (p.status) p.status[b[q]] = ok ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM;
        if (!ok && p.status_any) atomicOr(p.status_any, MBQC_STATUS_BAD_NORM);
        const double r = rsqrt(n2[q] * zn);
        const double ur = zr[q] * r, ui = zi[q] * r;  // unit phase / norm
        double2* o = stage_out ? (dyn + ((size_t)(threadIdx.x + q * kRegThreads) << k)) : (p.out + (b[q] << k));
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int d = p.tab.out_dst[i];
            if (d >= 0) o[d] = make_double2(re[q][i] * ur - im[q][i] * ui, re[q][i] * ui + im[q][i] * ur);
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
