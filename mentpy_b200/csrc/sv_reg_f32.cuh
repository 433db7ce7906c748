// complex64 mode of the register-resident batched state-vector kernel (window w <= 5).
//
// Same algorithm, plan tables and staging scheme as sv_reg.cuh (see there for the design notes);
// the state lives in fp32 registers, the pair update is 4 FFMA per pair on the full-rate FP32
// pipe, CZ signs are one LOP3 per value, outputs are complex64.  Angles are still read as fp64
// (the caller's dtype) and converted once per angle.  Accuracy target of the mode (BASELINE.json
// north_star): infidelity <= 1e-5 against the fp64 reference; measured ~1e-12..1e-11 on the
// benchmark patterns (tests/test_cuda_parity.py::test_complex64_mode).
#pragma once
#include "sv_reg.cuh"

namespace mbqc {

__device__ __forceinline__ float flip_sign_f(float x, uint32_t signbit) {
    return __int_as_float(__float_as_int(x) ^ (int)signbit);
}

template <int W, int S>
__device__ __forceinline__ void reg_stage_f(float (&re)[1 << W], float (&im)[1 << W], float c, float s,
                                            const uint32_t* __restrict__ sg) {
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
        const float tr = fmaf(c, re[j], fmaf(s, im[j], re[i]));
        const float ti = fmaf(c, im[j], fmaf(-s, re[j], im[i]));
        re[i] = tr;
        im[i] = ti;
        re[j] = flip_sign_f(tr, w[p]);
        im[j] = flip_sign_f(ti, w[p]);
        ++p;
    }
}

template <int W>
__device__ __forceinline__ void reg_step_any_f(float (&re)[1 << W], float (&im)[1 << W], int slot, float c,
                                               float s, const uint32_t* sg) {
    switch (slot) {
        case 0: reg_stage_f<W, 0>(re, im, c, s, sg); break;
        case 1: if constexpr (W > 1) reg_stage_f<W, 1>(re, im, c, s, sg); break;
        case 2: if constexpr (W > 2) reg_stage_f<W, 2>(re, im, c, s, sg); break;
        case 3: if constexpr (W > 3) reg_stage_f<W, 3>(re, im, c, s, sg); break;
        case 4: if constexpr (W > 4) reg_stage_f<W, 4>(re, im, c, s, sg); break;
        default: break;
    }
}

// DM = false: out [B][2^k] complex64; DM = true: out [B][2^k][2^k] complex64.
// The (cos, sin) tile re-uses the fp64 path's 16-byte slots: the raw fp64 angle is fetched into
// the low half, converted in place to a float2.
template <int W, bool DM>
__global__ void __launch_bounds__(128) sv_reg_kernel_f32(const __grid_constant__ SvRegParams pp, int staged) {
    constexpr int N = 1 << W;
    extern __shared__ double2 dyn[];
    const SvBatchParams& p = pp.base;
    const PlanTables& t = p.tab;
    const int T = t.n_angles, M = t.n_steps;
    const RegSmemLayout l = reg_smem_carve(dyn, M, pp.reg.sign_pitch, pp.reg.n_fixed);
    const int64_t b0 = (int64_t)blockIdx.x * kRegThreads;
    const int64_t b = b0 + threadIdx.x;
    const bool live = b < p.batch;
    const int k = t.n_out;
    const int samples = (int)min((int64_t)kRegThreads, p.batch - b0);
    stage_reg_tables(pp, l);
    if ((staged & 1) && live) fetch_own_row(p.angles + b * p.stride, l.cs + threadIdx.x, T, kRegThreads);
    cp_async_wait_all();
    __syncthreads();
    float re[N], im[N], zr = 1.f, zi = 0.f, n2 = 1.f;
    if (live) {
        double2* cs = l.cs + threadIdx.x;
        if (staged & 1) {
#pragma unroll 2
            for (int j = 0; j < T; ++j) {
                float sn, c;
                sincosf((float)cs[j * kRegThreads].x, &sn, &c);
                *reinterpret_cast<float2*>(&cs[j * kRegThreads]) = make_float2(c, sn);
            }
        }
        const float2* in = reinterpret_cast<const float2*>(p.inputs) + (p.input_mode == MBQC_INPUT_BATCH ? (b << t.n_in) : 0);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t sb = (t.init_sign << (31 - i)) & 0x80000000u;
            float2 v = make_float2((float)t.plus_amp, 0.f);
            if (p.input_mode != MBQC_INPUT_PLUS) {
                v = __ldg(in + t.init_src[i]);
                v.x *= (float)t.init_scale;
                v.y *= (float)t.init_scale;
            }
            re[i] = flip_sign_f(v.x, sb);
            im[i] = flip_sign_f(v.y, sb);
        }
        const double* row = p.angles + b * p.stride;
        for (int m = 0; m < M; ++m) {
            const uint32_t cw = l.cols[m];
            const int col = (int)(cw & 0xffffu);
            float c, s;
            if (col >= T) {
                const double2 f = l.fixed[col - T];
                c = (float)f.x;
                s = (float)f.y;
            } else if (staged & 1) {
                const float2 v = *reinterpret_cast<const float2*>(&cs[col * kRegThreads]);
                c = v.x;
                s = v.y;
            } else {
                sincosf((float)__ldg(row + col), &s, &c);
            }
            const float pr = 1.f + c;  // reference global phase factor (1 + e^{i theta})
            const float nzr = fmaf(zr, pr, -zi * s);
            zi = fmaf(zr, s, zi * pr);
            zr = nzr;
            reg_step_any_f<W>(re, im, (int)(cw >> 16), c, s, l.signs + m * pp.reg.sign_pitch);
            if ((m & 7) == 7) {  // keep magnitudes bounded (fp32 range)
                float a = 0.f;
#pragma unroll
                for (int i = 0; i < N; ++i) a = fmaf(re[i], re[i], fmaf(im[i], im[i], a));
                const float r = rsqrtf(a), rz = rsqrtf(zr * zr + zi * zi);
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    re[i] *= r;
                    im[i] *= r;
                }
                zr *= rz;
                zi *= rz;
            }
        }
        n2 = 0.f;
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (t.out_dst[i] >= 0) n2 = fmaf(re[i], re[i], fmaf(im[i], im[i], n2));
    }
    float2* stage = reinterpret_cast<float2*>(dyn);
    float2* outp = reinterpret_cast<float2*>(p.out);
    const bool stage_out = DM || (staged & 2);
    if (stage_out) __syncthreads();
    if (live) {
        const float zn = zr * zr + zi * zi;
        const bool ok = (n2 > 0.f) && (zn > 0.f) && isfinite(n2) && isfinite(zn);
        if (p.status) p.status[b] = ok ? MBQC_STATUS_OK : MBQC_STATUS_BAD_NORM;
        if (!ok && p.status_any) atomicOr(p.status_any, MBQC_STATUS_BAD_NORM);
        const float r = rsqrtf(n2) * rsqrtf(zn);
        const float ur = zr * r, ui = zi * r;
        float2* o = stage_out ? (stage + ((size_t)threadIdx.x << k)) : (outp + (b << k));
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int d = t.out_dst[i];
            if (d >= 0) o[d] = make_float2(re[i] * ur - im[i] * ui, re[i] * ui + im[i] * ur);
        }
    }
    if (!stage_out) return;
    __syncthreads();
    if constexpr (DM) {
        const int64_t total = (int64_t)samples << (2 * k);
        float2* o = outp + (b0 << (2 * k));
        const uint32_t km = (1u << k) - 1u;
        for (int64_t e = threadIdx.x; e < total; e += kRegThreads) {
            const float2* sv = stage + ((e >> (2 * k)) << k);
            const float2 x = sv[(e >> k) & km], y = sv[e & km];
            o[e] = make_float2(x.x * y.x + x.y * y.y, x.y * y.x - x.x * y.y);
        }
    } else {
        const int total = samples << k;
        float2* o = outp + (b0 << k);
        for (int e = threadIdx.x; e < total; e += kRegThreads) o[e] = stage[e];
    }
}

}  // namespace mbqc
