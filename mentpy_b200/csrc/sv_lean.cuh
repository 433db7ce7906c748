// Lean register-resident batched state-vector kernel (window w <= 5, reference schedules).
//
// Same contract as sv_reg_kernel (sv_reg.cuh): NumpySimulatorSV.run over a batch of angle sets,
// one thread per angle set, the 2^w amplitudes in registers
// (mentpy/simulators/np_simulator_sv.py:164-358, calculator/state_ops.py:42-74,
// operators/gates.py:62-72,127-143; per-measurement identity in common.cuh).
//
// Why a second kernel: ncu on sv_reg_kernel<3> (profiles/r01_*) showed 1638 issued instructions
// per warp for 10 measurements of which only ~530 are FP64 -- the FP64 pipe (2 cycles per warp
// instruction) was 45 % busy because two thirds of the issue slots went to address arithmetic,
// shared-memory table reads, moves and selects.  This kernel is built around the issue budget:
//   * the plan tables (angle column per step, CZ sign words per pair) live in the kernel
//     PARAMETER block: they are read through the uniform datapath (LDCU) and feed LOP3 as
//     uniform-register operands -- no shared-memory staging, no per-thread table loads;
//   * each warp fetches its 32 angle rows with ONE bulk async copy (cp.async.bulk, the TMA unit)
//     into shared memory, completion on a per-warp mbarrier: no per-thread cp.async, no CTA
//     barrier; a thread then reads one angle (LDS.64 at row + uniform offset) per measurement
//     and converts it just in time -- no (cos, sin) tile round trip through shared memory;
//   * sincos with absolute (not relative) accuracy: two-term Cody-Waite + plain Horner kernels,
//     19 FP64 instructions, no range branch (valid for |x| < 2^31; larger or non-finite angles
//     are reported through the status word);
//   * the tail of the pattern (measurements that append no qubit) and the output gather are
//     specialised at compile time on the single dead slot.
#pragma once
#include "common.cuh"

namespace mbqc {

constexpr int kLeanMaxSteps = 160;
constexpr int kLeanMaxSignWords = 640;
constexpr uint32_t kLeanFixedBit = 0x80000000u;

struct LeanParams {
    const double* __restrict__ angles;  // [B][T], contiguous rows, 16-byte aligned
    double2* __restrict__ out;
    int32_t* __restrict__ status;
    int32_t* __restrict__ status_any;
    const double2* __restrict__ inputs;
    const double2* __restrict__ fixed;  // (cos, sin) of fixed-angle steps
    int64_t batch;
    int32_t n_angles, n_steps, n_out, n_in;
    int32_t input_mode;
    int32_t n_full;     // steps [0, n_full) append a qubit, [n_full, n_steps) only project
    uint32_t init_sign;
    uint32_t pad0;
    double init_scale;
    uint8_t init_src[32];
    int8_t out_dst[32];
    uint32_t colofs[kLeanMaxSteps];     // byte offset of the step's angle inside a row | kLeanFixedBit | fixed index
    uint32_t signs[kLeanMaxSignWords];  // [M][2^(w-1)] sign words (0 / 0x80000000) of the pair partners
};

// ---- sincos with absolute accuracy ---------------------------------------------------------------
// x = q * pi/2 + r (two-term Cody-Waite: the product q * hi is exact inside the FMA, the dropped
// third term is < 2^31 * 1.5e-33), plain Horner kernels on |r| <= pi/4.  Absolute error <= 4e-16
// for |x| < 2^31 (tests/test_cuda_parity.py: sincos sweep); what the pattern needs is absolute
// accuracy of (cos, sin), not the library's relative accuracy near multiples of pi/2.
__device__ __forceinline__ void sincos_abs(double x, double& sn, double& cs) {
    const double magic = 6755399441055744.0;
    const double t = fma(x, kTrig[15], magic);
    const uint32_t q = (uint32_t)__double2loint(t);
    const double qd = t - magic;
    double r = fma(qd, -kTrig[12], x);
    r = fma(qd, -kTrig[13], r);
    const double z = r * r;
    double ps = fma(kTrig[5], z, kTrig[4]);
    double pc = fma(kTrig[11], z, kTrig[10]);
    ps = fma(ps, z, kTrig[3]);
    pc = fma(pc, z, kTrig[9]);
    ps = fma(ps, z, kTrig[2]);
    pc = fma(pc, z, kTrig[8]);
    ps = fma(ps, z, kTrig[1]);
    pc = fma(pc, z, kTrig[7]);
    ps = fma(ps, z, kTrig[0]);
    pc = fma(pc, z, kTrig[6]);
    const double s0 = fma(r * z, ps, r);
    const double c0 = fma(z, fma(z, pc, -0.5), 1.0);
    const bool odd = (q & 1u) != 0u;
    const double sa = odd ? c0 : s0;
    const double ca = odd ? s0 : c0;
    const uint32_t b1 = (q << 30) & 0x80000000u;  // bit 1 of q
    sn = flip_sign(sa, b1);
    cs = flip_sign(ca, b1 ^ (q << 31));  // bit 1 of q + 1
}

// Table-driven variant: x = k * (2 pi / 128) + r with |r| <= pi / 128, degree-7 / degree-6 Taylor
// kernels for (sin r, cos r) (truncation < 1e-20 / 4e-18) and one rotation by the tabulated
// (cos, sin)(k * 2 pi / 128): 16 FP64 instructions and NO quadrant selects or sign fix-ups --
// one 16-byte shared-memory gather instead of ~10 ALU instructions.  `tab` is the CTA's
// shared-memory copy of kTrigTable128.
#include "trig_table128.inc"
static __device__ const double2 kTrigTable128[MBQC_TRIG128_N] = {MBQC_TRIG128_TABLE_ROWS};

__device__ __forceinline__ void sincos_tab128(double x, double& sn, double& cs, const double2* __restrict__ tab) {
    const double magic = 6755399441055744.0;
    const double t = fma(x, MBQC_TRIG128_INV, magic);
    const uint32_t k = (uint32_t)__double2loint(t) & (MBQC_TRIG128_N - 1);
    const double kd = t - magic;
    double r = fma(kd, -MBQC_TRIG128_C1, x);
    r = fma(kd, -MBQC_TRIG128_C2, r);
    const double2 ck = tab[k];
    const double z = r * r;
    double ps = fma(-1.9841269841269841e-04, z, 8.3333333333333333e-03);  // -1/7!, 1/5!
    double pc = fma(-1.3888888888888889e-03, z, 4.1666666666666664e-02);  // -1/6!, 1/4!
    ps = fma(ps, z, -1.6666666666666666e-01);
    pc = fma(pc, z, -0.5);
    const double sr = fma(r * z, ps, r);
    const double cr = fma(pc, z, 1.0);
    cs = fma(ck.x, cr, -(ck.y * sr));
    sn = fma(ck.y, cr, ck.x * sr);
}

#ifndef MBQC_LEAN_SINCOS_TAB
#define MBQC_LEAN_SINCOS_TAB 1
#endif

// ---- bulk copy + mbarrier (one per warp) ---------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LEAN_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LEAN_DONE;\n"
        "bra LEAN_WAIT;\n"
        "LEAN_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// ---- one measurement -----------------------------------------------------------------------------
// APPEND: the measured slot is recycled for a fresh |+> qubit (partner amplitudes = +-t);
// otherwise (tail of the pattern) the slot dies and only the bit-clear half stays meaningful.
template <int W, int S, bool APPEND>
__device__ __forceinline__ void lean_stage(double (&re)[1 << W], double (&im)[1 << W], double c, double s,
                                           const uint32_t* __restrict__ sg) {
    int pidx = 0;
#pragma unroll
    for (int i = 0; i < (1 << W); ++i) {
        if (i & (1 << S)) continue;
        const int j = i | (1 << S);
        const double tr = fma(c, re[j], fma(s, im[j], re[i]));
        const double ti = fma(c, im[j], fma(-s, re[j], im[i]));
        re[i] = tr;
        im[i] = ti;
        if constexpr (APPEND) {
            const uint32_t w = sg[pidx];
            re[j] = flip_sign(tr, w);
            im[j] = flip_sign(ti, w);
        }
        ++pidx;
    }
}

// kernel-side view of one thread's angle row + the plan tables in the parameter block
template <bool FIXED>
__device__ __forceinline__ void lean_angle(const LeanParams& p, const char* row, const double2* trig, int m,
                                           double& c, double& s, uint32_t& big) {
    const uint32_t co = p.colofs[m];
    if constexpr (FIXED) {
        if (co & kLeanFixedBit) {
            const double2 f = __ldg(p.fixed + (co & ~kLeanFixedBit));
            c = f.x;
            s = f.y;
            return;
        }
    }
    const double th = *reinterpret_cast<const double*>(row + co);
    big = max(big, (uint32_t)__double2hiint(th) & 0x7fffffffu);
#if MBQC_LEAN_SINCOS_TAB
    sincos_tab128(th, s, c, trig);
#else
    sincos_abs(th, s, c);
#endif
}

template <int W, int S, bool APPEND, bool FIXED, bool PHASE>
__device__ __forceinline__ void lean_step(const LeanParams& p, const char* row, const double2* trig, int m,
                                          double (&re)[1 << W], double (&im)[1 << W], double& zr, double& zi,
                                          uint32_t& big) {
    double c, s;
    lean_angle<FIXED>(p, row, trig, m, c, s, big);
    if constexpr (PHASE) {  // (zr, zi) *= (1 + c, s)
        const double nzr = fma(-zi, s, fma(zr, c, zr));
        zi = fma(zr, s, fma(zi, c, zi));
        zr = nzr;
    }
    lean_stage<W, S, APPEND>(re, im, c, s, p.signs + m * (1 << (W - 1)));
}

template <int W>
__device__ __forceinline__ void lean_renorm(double (&re)[1 << W], double (&im)[1 << W], double& zr, double& zi) {
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

// OUT: 0 = [B][2^k] amplitudes stored straight from registers; 1 = the same, collected in shared
// memory and written CTA-coalesced (output buffers in page-locked host memory); 2 = [B][4^k]
// |psi><psi| (np_simulator_sv.py:292-293), CTA-coalesced.
constexpr int kLeanOutDirect = 0, kLeanOutStaged = 1, kLeanOutDM = 2;

template <int W, int CTA>
struct LeanTraits {
    static constexpr int kMinBlocks = (W <= 3) ? (1024 / CTA) : (W == 4 ? 640 / CTA : 384 / CTA);
};

template <int W, int CTA, bool FIXED, int OUT>
__global__ void __launch_bounds__(CTA, LeanTraits<W, CTA>::kMinBlocks)
sv_lean_kernel(const __grid_constant__ LeanParams p) {
    constexpr int N = 1 << W;
    constexpr bool PHASE = OUT != kLeanOutDM;
    extern __shared__ __align__(16) unsigned char lean_smem[];
    __shared__ __align__(8) uint64_t bars[CTA / 32];
    __shared__ __align__(16) double2 s_trig[MBQC_TRIG128_N];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = p.n_angles, M = p.n_steps;
    const int64_t b0 = (int64_t)blockIdx.x * CTA;
    const int64_t wb0 = b0 + warp * 32;
    const int64_t b = wb0 + lane;
    const bool live = b < p.batch;
    const int rows = (int)max((int64_t)0, min((int64_t)32, p.batch - wb0));
    const uint32_t bytes = (uint32_t)(rows * T) * 8u;
    const uint32_t bulk = bytes & ~15u;
    double* wtile = reinterpret_cast<double*>(lean_smem) + (size_t)warp * 32 * T;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[warp]);
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (bulk) {
            mbar_expect_tx(bar, bulk);
            bulk_g2s((uint32_t)__cvta_generic_to_shared(wtile), p.angles + wb0 * T, bulk, bar);
        }
        if (bytes & 8u) wtile[rows * T - 1] = __ldg(p.angles + wb0 * T + rows * T - 1);
    }
    __syncwarp();
#if MBQC_LEAN_SINCOS_TAB
    for (int i = tid; i < MBQC_TRIG128_N; i += CTA) s_trig[i] = kTrigTable128[i];
    __syncthreads();  // early: every warp gets here before its angle rows have arrived
#endif

    double re[N], im[N], zr = 1.0, zi = 0.0;
    if (p.input_mode == MBQC_INPUT_PLUS) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            re[i] = flip_sign(1.0, (p.init_sign << (31 - i)) & 0x80000000u);
            im[i] = 0.0;
        }
    } else {
        const double2* in = p.inputs + (p.input_mode == MBQC_INPUT_BATCH ? ((live ? b : 0) << p.n_in) : 0);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double2 v = __ldg(in + p.init_src[i]);
            const uint32_t sb = (p.init_sign << (31 - i)) & 0x80000000u;
            re[i] = flip_sign(v.x, sb);
            im[i] = flip_sign(v.y, sb);
        }
    }
    if (bulk) mbar_wait(bar, 0);
    // dead lanes of a partial warp read row 0 of the warp (valid data, results unused)
    const char* row = reinterpret_cast<const char*>(wtile + (size_t)(live ? lane : 0) * T);
    uint32_t big = 0;
    const int n_full = p.n_full;
    int m = 0;
    for (; m + W <= n_full; m += W) {
        static_for<W>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            lean_step<W, W - 1 - u, true, FIXED, PHASE>(p, row, s_trig, m + u, re, im, zr, zi, big);
        });
        if (((m + W) >> 5) != (m >> 5)) lean_renorm<W>(re, im, zr, zi);  // long patterns: keep magnitudes bounded
    }
    // m is a multiple of W: the remaining appended steps and the tail, slots still compile-time
    static_for<W>([&](auto uc) {
        constexpr int u = decltype(uc)::value;
        if (m + u < M) {
            if (m + u < n_full) lean_step<W, W - 1 - u, true, FIXED, PHASE>(p, row, s_trig, m + u, re, im, zr, zi, big);
            else lean_step<W, W - 1 - u, false, FIXED, PHASE>(p, row, s_trig, m + u, re, im, zr, zi, big);
        }
    });
    if (m + W < M) {  // more than W leftover steps: window with several dead slots
        static_for<W>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            if (m + W + u < M) lean_step<W, W - 1 - u, false, FIXED, PHASE>(p, row, s_trig, m + W + u, re, im, zr, zi, big);
        });
    }

    // ---- output ----
    const int k = p.n_out;
    double n2 = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (p.out_dst[i] >= 0) n2 = fma(re[i], re[i], fma(im[i], im[i], n2));
    double ur, ui;
    if constexpr (PHASE) {
        const double zn = zr * zr + zi * zi;
        const double x = n2 * zn;
        const double r = rsqrt(x);
        ur = zr * r;
        ui = zi * r;
        n2 = x;
    } else {
        ur = rsqrt(n2);
        ui = 0.0;
    }
    // ok <=> n2 finite and > 0 (integer test on the exponent field) and every angle in range
    const uint32_t hx = (uint32_t)__double2hiint(n2);
    const bool ok = (hx - 0x00100000u < 0x7fe00000u) && (big < 0x41e00000u);
    if (live) {
        if (p.status) p.status[b] = ok ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM;
        if (!ok && p.status_any) atomicOr(p.status_any, MBQC_STATUS_BAD_NORM);
    }
    if constexpr (OUT == kLeanOutDirect) {
        if (!live) return;
        double2* o = p.out + (b << k);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int d = p.out_dst[i];
            if (d >= 0) {
                if constexpr (PHASE) o[d] = make_double2(re[i] * ur - im[i] * ui, fma(re[i], ui, im[i] * ur));
                else o[d] = make_double2(re[i] * ur, im[i] * ur);
            }
        }
    } else {
        __syncthreads();  // every warp is done with its angle rows: the buffer becomes the output stage
        double2* stage = reinterpret_cast<double2*>(lean_smem);
        if (live) {
            double2* o = stage + ((size_t)tid << k);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int d = p.out_dst[i];
                if (d >= 0) {
                    if constexpr (PHASE) o[d] = make_double2(re[i] * ur - im[i] * ui, fma(re[i], ui, im[i] * ur));
                    else o[d] = make_double2(re[i] * ur, im[i] * ur);
                }
            }
        }
        __syncthreads();
        const int samples = (int)min((int64_t)CTA, p.batch - b0);
        if constexpr (OUT == kLeanOutDM) {
            const int64_t total = (int64_t)samples << (2 * k);
            double2* o = p.out + (b0 << (2 * k));
            const uint32_t km = (1u << k) - 1u;
            for (int64_t e = tid; e < total; e += CTA) {
                const double2* sv = stage + ((e >> (2 * k)) << k);
                const double2 x = sv[(e >> k) & km], y = sv[e & km];
                o[e] = make_double2(x.x * y.x + x.y * y.y, x.y * y.x - x.x * y.y);
            }
        } else {
            const int total = samples << k;
            double2* o = p.out + (b0 << k);
            for (int e = tid; e < total; e += CTA) o[e] = stage[e];
        }
    }
}

}  // namespace mbqc
