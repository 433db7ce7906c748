// C ABI of the MBQC hot path (include/mbqc_b200.h): plan lowering tables + kernel launches.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "common.cuh"
#include "dm_batch.cuh"
#include "dm_reg.cuh"
#include "sample.cuh"
#include "grad_batch.cuh"
#include "sv_batch.cuh"
#include "sv_reg.cuh"
#include "sv_reg_f32.cuh"
#include "stream.cuh"
#include "calc.cuh"
#include <vector>

#include "host_util.h"

using namespace mbqc;

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(MBQC_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CUDA_TRY(x)                                        \
    do {                                                   \
        cudaError_t _e = (x);                              \
        if (_e != cudaSuccess) return cuda_fail(_e, #x);   \
    } while (0)

int after_launch(const char* name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    return MBQC_OK;
}

}  // namespace

// helpers shared with the other translation units of the library (host_util.h)
int mbqc_set_error(int code, const char* msg) { return fail(code, "%s", msg); }
int mbqc_cuda_error(cudaError_t e, const char* what) { return cuda_fail(e, what); }
int mbqc_after_launch(const char* name) { return after_launch(name); }

namespace {
int ilog2_ceil(unsigned v) {
    int l = 0;
    while ((1u << l) < v) ++l;
    return l;
}
}  // namespace

extern "C" {

const char* mbqc_last_error(void) { return g_err; }
const char* mbqc_version(void) { return "mentpy_b200 0.1 (sm_100a)"; }
int64_t mbqc_launch_count(void) { return (int64_t)g_launches.load(); }

static int plan_create_impl(const mbqc_step* steps, int32_t n_steps, int32_t window, int32_t n_inputs,
                            int32_t n_outputs, int32_t n_angles, const int32_t* input_slot,
                            const uint64_t* init_cz_mask, const int32_t* output_slot,
                            const mbqc_noise* noise, mbqc_plan** out, bool upload) {
    if (!out) return fail(MBQC_E_ARG, "out is NULL");
    *out = nullptr;
    if (window < 1 || window > MBQC_MAX_WINDOW) return fail(MBQC_E_ARG, "window %d out of range [1,%d]", window, MBQC_MAX_WINDOW);
    if (n_steps < 0 || (n_steps > 0 && !steps)) return fail(MBQC_E_ARG, "bad step list");
    if (n_inputs < 0 || n_inputs > window || n_inputs > kMaxIO) return fail(MBQC_E_ARG, "n_inputs %d not in [0, min(window,%d)]", n_inputs, kMaxIO);
    if (n_outputs < 0 || n_outputs > window || n_outputs > kMaxIO) return fail(MBQC_E_ARG, "n_outputs %d not in [0, min(window,%d)]", n_outputs, kMaxIO);
    if (n_angles < 0) return fail(MBQC_E_ARG, "n_angles < 0");
    if ((n_inputs && !input_slot) || (n_outputs && !output_slot) || !init_cz_mask) return fail(MBQC_E_ARG, "NULL slot table");
    const uint64_t wmask = (window >= 64) ? ~0ull : ((1ull << window) - 1ull);
    uint64_t seen = 0;
    for (int q = 0; q < n_inputs; ++q) {
        if (input_slot[q] < 0 || input_slot[q] >= window || ((seen >> input_slot[q]) & 1)) return fail(MBQC_E_ARG, "input_slot[%d] invalid", q);
        seen |= 1ull << input_slot[q];
    }
    seen = 0;
    for (int q = 0; q < n_outputs; ++q) {
        if (output_slot[q] < 0 || output_slot[q] >= window || ((seen >> output_slot[q]) & 1)) return fail(MBQC_E_ARG, "output_slot[%d] invalid", q);
        seen |= 1ull << output_slot[q];
    }
    for (int m = 0; m < n_steps; ++m) {
        const mbqc_step& s = steps[m];
        if (s.slot < 0 || s.slot >= window) return fail(MBQC_E_ARG, "step %d: slot %d outside window", m, s.slot);
        if (s.angle_idx < -1 || s.angle_idx >= n_angles) return fail(MBQC_E_ARG, "step %d: angle_idx %d outside [ -1, %d)", m, s.angle_idx, n_angles);
        if (s.plane < MBQC_PLANE_XY || s.plane > MBQC_PLANE_XYZ) return fail(MBQC_E_ARG, "step %d: plane %d unknown", m, s.plane);
        if (s.plane == MBQC_PLANE_XYZ && s.angle_idx >= 0) return fail(MBQC_E_ARG, "step %d: plane XYZ takes two fixed angles (ment.py:239-251), not a column of the angle matrix", m);
        if (s.cond_mask) {
            if (s.alt_plane < MBQC_PLANE_XY || s.alt_plane > MBQC_PLANE_XYZ || s.alt_plane == MBQC_PLANE_Z || s.plane == MBQC_PLANE_Z)
                return fail(MBQC_E_ARG, "step %d: controlled measurements take the planes XY, XZ, YZ, XYZ", m);
            if (s.alt_angle_idx < -1 || s.alt_angle_idx >= n_angles) return fail(MBQC_E_ARG, "step %d: alt_angle_idx %d outside [-1, %d)", m, s.alt_angle_idx, n_angles);
            if (s.alt_plane == MBQC_PLANE_XYZ && s.alt_angle_idx >= 0) return fail(MBQC_E_ARG, "step %d: plane XYZ takes two fixed angles", m);
            if (m < 32 && (s.cond_mask >> m)) return fail(MBQC_E_ARG, "step %d: condition reads an outcome before the first measurement", m);
            if (__builtin_popcount(s.cond_mask) > 5) return fail(MBQC_E_ARG, "step %d: a condition reads at most 5 outcomes (got %d)", m, __builtin_popcount(s.cond_mask));
        }
        if (s.nbr_mask & ~wmask) return fail(MBQC_E_ARG, "step %d: nbr_mask outside window", m);
        if ((s.nbr_mask >> s.slot) & 1ull) return fail(MBQC_E_ARG, "step %d: nbr_mask contains the step's own slot", m);
    }
    mbqc_plan* pl = new (std::nothrow) mbqc_plan();
    if (!pl) return fail(MBQC_E_CUDA, "out of host memory");
    memset(&pl->tab, 0, sizeof(pl->tab));
    PlanTables& t = pl->tab;
    t.window = window;
    t.n_steps = n_steps;
    t.n_in = n_inputs;
    t.n_out = n_outputs;
    t.n_angles = n_angles;
    t.has_noise = noise ? 1 : 0;
    if (noise) t.noise = *noise;
    t.init_scale = std::exp2(-0.5 * (window - n_inputs));
    t.plus_amp = std::exp2(-0.5 * window);
    for (int q = 0; q < n_inputs; ++q) t.in_slot[q] = input_slot[q];
    for (int q = 0; q < n_outputs; ++q) t.out_slot[q] = output_slot[q];
    if (window <= kMaxSlotsSmall)
        for (int a = 0; a < window; ++a) t.init_cz[a] = init_cz_mask[a] & wmask;
    if (window <= MBQC_MAX_WINDOW_REG) {
        uint64_t outmask = 0;
        for (int q = 0; q < n_outputs; ++q) outmask |= 1ull << output_slot[q];
        for (uint32_t i = 0; i < (1u << window); ++i) {
            t.init_src[i] = (uint8_t)init_source_index(t, i);
            if (init_sign_bit(t, i)) t.init_sign |= 1u << i;
            if (i & ~outmask) {
                t.out_dst[i] = -1;
            } else {
                uint32_t d = 0;
                for (int q = 0; q < n_outputs; ++q) d |= ((i >> output_slot[q]) & 1u) << (n_outputs - 1 - q);
                t.out_dst[i] = (int8_t)d;
            }
        }
    }
    pl->h_steps = new (std::nothrow) StepDev[n_steps > 0 ? n_steps : 1];
    if (!pl->h_steps) {
        delete pl;
        return fail(MBQC_E_CUDA, "out of host memory");
    }
    for (int m = 0; m < n_steps; ++m) {
        StepDev& d = pl->h_steps[m];
        const mbqc_step& s = steps[m];
        d.slot = s.slot;
        d.angle_idx = s.angle_idx;
        d.plane = s.plane;
        d.flags = s.flags;
        d.fc = s.fixed_cos;
        d.fs = s.fixed_sin;
        d.nbr_mask = (s.flags & MBQC_STEP_APPEND) ? s.nbr_mask : 0ull;
        d.flipmask = 0;
        d.pad = 0;
        d.fz = s.fixed_z;
        d.cond_mask = s.cond_mask;
        d.cond_table = s.cond_mask ? s.cond_table : 0u;
        d.alt_plane = s.cond_mask ? s.alt_plane : s.plane;
        d.alt_angle_idx = s.cond_mask ? s.alt_angle_idx : -1;
        d.afc = s.alt_cos;
        d.afs = s.alt_sin;
        d.afz = s.alt_z;
        if (window <= MBQC_MAX_WINDOW_REG)
            for (uint32_t i = 0; i < (1u << window); ++i)
                if (((i >> s.slot) & 1u) && parity64((uint64_t)i & d.nbr_mask)) d.flipmask |= 1u << i;
    }
    pl->d_steps = nullptr;
    pl->device = -1;
    cudaError_t e = cudaSuccess;
    if (upload) e = cudaGetDevice(&pl->device);
    if (upload && e == cudaSuccess) e = cudaMalloc(&pl->d_steps, sizeof(StepDev) * (n_steps > 0 ? n_steps : 1));
    if (upload && e == cudaSuccess && n_steps > 0)
        e = cudaMemcpy(pl->d_steps, pl->h_steps, sizeof(StepDev) * n_steps, cudaMemcpyHostToDevice);
    pl->d_reg_blob = nullptr;
    pl->d_reg_cols = pl->d_reg_signs = nullptr;
    pl->d_reg_fixed = nullptr;
    pl->reg_n_fixed = pl->reg_sign_pitch = pl->reg_periodic = 0;
    if (e == cudaSuccess && window <= MBQC_MAX_WINDOW_REG) {
        // register-kernel tables: packed (slot<<16 | column), sign words per pair, fixed (cos,sin)
        const int np = 1 << (window - 1);
        const int sp = np < 4 ? 4 : np;
        const int mpad = (n_steps + 3) & ~3;
        std::vector<uint32_t> signs((size_t)(n_steps > 0 ? n_steps : 1) * sp, 0u), cols(mpad > 0 ? mpad : 4, 0u);
        std::vector<double2> fixed;
        int periodic = 1;
        for (int m = 0; m < n_steps; ++m) {
            const StepDev& d = pl->h_steps[m];
            if (d.slot != window - 1 - (m % window)) periodic = 0;
            uint32_t col;
            if (d.angle_idx >= 0) {
                col = (uint32_t)d.angle_idx;
            } else {
                col = (uint32_t)(n_angles + (int)fixed.size());
                fixed.push_back(make_double2(d.fc, d.fs));
            }
            cols[m] = ((uint32_t)d.slot << 16) | (col & 0xffffu);
            int pidx = 0;
            for (uint32_t i = 0; i < (1u << window); ++i) {
                if ((i >> d.slot) & 1u) continue;
                const uint32_t j = i | (1u << d.slot);
                signs[(size_t)m * sp + pidx] = ((d.flipmask >> j) & 1u) ? 0x80000000u : 0u;
                ++pidx;
            }
        }
        if (n_angles + (int)fixed.size() > 0xffff) {
            // the packed (slot << 16 | column) words of the register kernels cannot name the column
            delete[] pl->h_steps;
            delete pl;
            return fail(MBQC_E_UNSUPPORTED, "window <= %d patterns are limited to 65535 angle columns (got %d)",
                        MBQC_MAX_WINDOW_REG, n_angles + (int)fixed.size());
        }
        const size_t b_signs = signs.size() * 4, b_cols = cols.size() * 4, b_fixed = (fixed.size() ? fixed.size() : 1) * 16;
        pl->reg_n_fixed = (int)fixed.size();
        pl->reg_sign_pitch = sp;
        pl->reg_periodic = periodic;
        if (upload) e = cudaMalloc(&pl->d_reg_blob, b_signs + b_cols + b_fixed);
        if (upload && e == cudaSuccess) {
            char* base = (char*)pl->d_reg_blob;
            e = cudaMemcpy(base, signs.data(), b_signs, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(base + b_signs, cols.data(), b_cols, cudaMemcpyHostToDevice);
            if (e == cudaSuccess && !fixed.empty()) e = cudaMemcpy(base + b_signs + b_cols, fixed.data(), fixed.size() * 16, cudaMemcpyHostToDevice);
            pl->d_reg_signs = (const uint32_t*)base;
            pl->d_reg_cols = (const uint32_t*)(base + b_signs);
            pl->d_reg_fixed = (const double2*)(base + b_signs + b_cols);
        }
    }
    if (e != cudaSuccess) {
        if (pl->d_steps) cudaFree(pl->d_steps);
        if (pl->d_reg_blob) cudaFree(pl->d_reg_blob);
        delete[] pl->h_steps;
        delete pl;
        return cuda_fail(e, "plan upload");
    }
    pl->lean = nullptr;
    pl->lean_fixed = 0;
    pl->jit = nullptr;
    mbqc_lean_build_proto(pl);
    *out = pl;
    return MBQC_OK;
}

int mbqc_plan_create(const mbqc_step* steps, int32_t n_steps, int32_t window, int32_t n_inputs,
                     int32_t n_outputs, int32_t n_angles, const int32_t* input_slot,
                     const uint64_t* init_cz_mask, const int32_t* output_slot,
                     const mbqc_noise* noise, mbqc_plan** out) {
    return plan_create_impl(steps, n_steps, window, n_inputs, n_outputs, n_angles, input_slot, init_cz_mask, output_slot,
                            noise, out, true);
}

int mbqc_plan_create_hostonly(const mbqc_step* steps, int32_t n_steps, int32_t window, int32_t n_inputs,
                              int32_t n_outputs, int32_t n_angles, const int32_t* input_slot,
                              const uint64_t* init_cz_mask, const int32_t* output_slot,
                              const mbqc_noise* noise, mbqc_plan** out) {
    return plan_create_impl(steps, n_steps, window, n_inputs, n_outputs, n_angles, input_slot, init_cz_mask, output_slot,
                            noise, out, false);
}

void mbqc_plan_destroy(mbqc_plan* plan) {
    if (!plan) return;
    if (plan->d_steps) cudaFree(plan->d_steps);
    if (plan->d_reg_blob) cudaFree(plan->d_reg_blob);
    if (plan->d_ff) cudaFree(plan->d_ff);
    mbqc_jit_free(plan);
    mbqc_lean_free_proto(plan);
    delete[] plan->h_steps;
    delete plan;
}

int mbqc_plan_set_feedforward(mbqc_plan* plan, const mbqc_feedforward* ff, int32_t n_steps) {
    if (!plan) return fail(MBQC_E_ARG, "plan is NULL");
    if (n_steps != plan->tab.n_steps) return fail(MBQC_E_ARG, "n_steps %d != plan steps %d", n_steps, plan->tab.n_steps);
    if (n_steps > 0 && !ff) return fail(MBQC_E_ARG, "ff is NULL");
    const int k = plan->tab.n_out;
    for (int m = 0; m < n_steps; ++m) {
        if ((k < 32 && ((ff[m].outx | ff[m].outz) >> k)) || k > 16)
            return fail(MBQC_E_ARG, "step %d: byproduct mask names an output >= %d (max 16 outputs)", m, k);
        if (m < 32 && (((uint64_t)ff[m].xdep | ff[m].zdep) >> m))
            return fail(MBQC_E_ARG, "step %d depends on an outcome before the first measurement", m);
    }
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(cudaSetDevice(plan->device));
    if (plan->d_ff) cudaFree(plan->d_ff);
    plan->d_ff = nullptr;
    static_assert(sizeof(mbqc_feedforward) == sizeof(mbqc::FeedForwardDev), "feed-forward record layout");
    if (n_steps > 0) {
        CUDA_TRY(cudaMalloc(&plan->d_ff, sizeof(mbqc_feedforward) * n_steps));
        CUDA_TRY(cudaMemcpy(plan->d_ff, ff, sizeof(mbqc_feedforward) * n_steps, cudaMemcpyHostToDevice));
    }
    cudaSetDevice(prev);
    return MBQC_OK;
}

int32_t mbqc_plan_window(const mbqc_plan* plan) { return plan ? plan->tab.window : -1; }
int32_t mbqc_plan_num_steps(const mbqc_plan* plan) { return plan ? plan->tab.n_steps : -1; }
int32_t mbqc_plan_num_outputs(const mbqc_plan* plan) { return plan ? plan->tab.n_out : -1; }

}  // extern "C"

static int check_batch_args(const mbqc_plan* plan, const double* d_angles, int64_t stride,
                            const void* d_inputs, int32_t input_mode, int64_t batch, void* d_out) {
    if (!plan) return fail(MBQC_E_ARG, "plan is NULL");
    if (plan->device < 0) return fail(MBQC_E_ARG, "host-only plan (mbqc_plan_create_hostonly) cannot be run");
    if (batch < 0) return fail(MBQC_E_ARG, "batch < 0");
    if (plan->tab.n_angles > 0 && !d_angles) return fail(MBQC_E_ARG, "d_angles is NULL");
    if (stride < plan->tab.n_angles) return fail(MBQC_E_ARG, "angle_stride %lld < n_angles %d", (long long)stride, plan->tab.n_angles);
    if (input_mode < MBQC_INPUT_PLUS || input_mode > MBQC_INPUT_BATCH) return fail(MBQC_E_ARG, "input_mode %d unknown", input_mode);
    if (input_mode != MBQC_INPUT_PLUS && !d_inputs) return fail(MBQC_E_ARG, "d_inputs is NULL");
    if (!d_out) return fail(MBQC_E_ARG, "d_out is NULL");
    return MBQC_OK;
}

static void fill_sv_params(SvBatchParams& p, const mbqc_plan* plan, const double* d_angles,
                           int64_t stride, const void* d_inputs, int32_t input_mode, int64_t batch,
                           void* d_out, int32_t* d_status) {
    memset(&p, 0, sizeof(p));
    p.tab = plan->tab;
    p.steps = plan->d_steps;
    p.angles = d_angles;
    p.stride = stride;
    p.inputs = (const double2*)d_inputs;
    p.input_mode = input_mode;
    p.batch = batch;
    p.out = (double2*)d_out;
    p.status = d_status;
}

static void fill_reg_params(SvRegParams& rp, const SvBatchParams& p, const mbqc_plan* plan) {
    rp.base = p;
    rp.reg.cols = plan->d_reg_cols;
    rp.reg.signs = plan->d_reg_signs;
    rp.reg.fixed = plan->d_reg_fixed;
    rp.reg.n_fixed = plan->reg_n_fixed;
    rp.reg.sign_pitch = plan->reg_sign_pitch;
    rp.reg.periodic = plan->reg_periodic;
}

// dynamic shared memory of the register kernels: plan tables + (staged) the CTA's (cos, sin)
// tile and raw angle tile; staging is skipped when the tile does not fit the budget
template <int W, bool DM>
static int launch_sv_reg_w(const SvBatchParams& p, const mbqc_plan* plan, cudaStream_t st, bool coalesced_out) {
    constexpr int spt = RegKernelTraits<W>::kSPT;
    const int threads = kRegThreads;
    const int per_cta = threads * spt;
    const unsigned blocks = (unsigned)((p.batch + per_cta - 1) / per_cta);
    SvRegParams rp;
    fill_reg_params(rp, p, plan);
    const size_t tables = reg_smem_tables_bytes(p.tab.n_steps, rp.reg.sign_pitch, rp.reg.n_fixed);
    if (tables > 200 * 1024)
        return fail(MBQC_E_UNSUPPORTED, "%d measurements at window %d: the step tables (%zu bytes) exceed the shared-memory budget of the register kernels",
                    p.tab.n_steps, W, tables);
    const size_t tile = (size_t)per_cta * p.tab.n_angles * sizeof(double2);
    const int staged = (p.tab.n_angles > 0 && tables + tile <= 100 * 1024) ? 1 : 0;
    size_t smem = tables + (staged ? tile : 0);
    const size_t stage = ((size_t)per_cta << p.tab.n_out) * sizeof(double2);  // output stage re-uses the buffer
    if ((DM || coalesced_out) && stage > smem) smem = stage;
    auto kern = sv_reg_kernel<W, DM>;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, threads, smem, st>>>(rp, staged | ((DM || coalesced_out) ? 2 : 0));
    return after_launch("sv_reg_kernel");
}

template <bool DM>
static int launch_sv_reg(const SvBatchParams& p, const mbqc_plan* plan, cudaStream_t st, bool coalesced_out) {
    switch (p.tab.window) {
        case 1: return launch_sv_reg_w<1, DM>(p, plan, st, coalesced_out);
        case 2: return launch_sv_reg_w<2, DM>(p, plan, st, coalesced_out);
        case 3: return launch_sv_reg_w<3, DM>(p, plan, st, coalesced_out);
        case 4: return launch_sv_reg_w<4, DM>(p, plan, st, coalesced_out);
        default: return launch_sv_reg_w<5, DM>(p, plan, st, coalesced_out);
    }
}

// call_batch: size of the API call this launch is a piece of (the host pipeline launches chunks);
// the choice of kernel depends on the call, not on how it was cut up.
static int launch_sv(const SvBatchParams& p, const mbqc_plan* plan, int out_form, cudaStream_t st, bool coalesced_out = false,
                     int64_t call_batch = -1) {
    const int w = p.tab.window;
    if (w <= MBQC_MAX_WINDOW_REG) {
        int rc = 0;
        const int out_mode = out_form == MBQC_OUT_DM ? MBQC_LEAN_OUT_DM : (coalesced_out ? MBQC_LEAN_OUT_STAGED : MBQC_LEAN_OUT_DIRECT);
        if (mbqc_jit_try_launch(p, plan, out_mode, st, call_batch < 0 ? p.batch : call_batch, &rc)) return rc;
        if (mbqc_lean_try_launch(p, plan, out_mode, st, &rc)) return rc;
    }
    if (w <= MBQC_MAX_WINDOW_REG)
        return out_form == MBQC_OUT_DM ? launch_sv_reg<true>(p, plan, st, true) : launch_sv_reg<false>(p, plan, st, coalesced_out);
    if (w > MBQC_MAX_WINDOW_SMEM_SV)
        return fail(MBQC_E_UNSUPPORTED, "batched SV covers window <= %d (got %d); use the streaming calls", MBQC_MAX_WINDOW_SMEM_SV, w);
    // threads per sample: one per group of 8 amplitudes (three measurements per pass) up to 256
    int tps_log2 = w - 3;
    if (tps_log2 > 8) tps_log2 = 8;
    const int tps = 1 << tps_log2;
    int spb = 256 / tps;
    if (spb < 1) spb = 1;
    const size_t smem = (size_t)spb * ((16ull << w) + 16ull * tps);  // amplitudes + a strip of staged (cos, sin)
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(sv_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)((p.batch + spb - 1) / spb);
    sv_smem_kernel<<<blocks, tps * spb, smem, st>>>(p, tps_log2, spb, out_form == MBQC_OUT_DM ? 1 : 0);
    return after_launch("sv_smem_kernel");
}

extern "C" {

int mbqc_run_batch_sv(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                      const void* d_inputs, int32_t input_mode, int64_t batch, void* d_out,
                      int32_t out_form, int32_t* d_status, void* stream) {
    int rc = check_batch_args(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out);
    if (rc) return rc;
    if (out_form != MBQC_OUT_SV && out_form != MBQC_OUT_DM) return fail(MBQC_E_ARG, "out_form %d unknown", out_form);
    for (int m = 0; m < plan->tab.n_steps; ++m)
        if (plan->h_steps[m].plane != MBQC_PLANE_XY || plan->h_steps[m].cond_mask)
            return fail(MBQC_E_ARG, "step %d: only the XY plane is supported on the state-vector path", m);
    if (batch == 0) return MBQC_OK;
    SvBatchParams p;
    fill_sv_params(p, plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out, d_status);
    return launch_sv(p, plan, out_form, (cudaStream_t)stream);
}

}  // extern "C"

// ---- host-buffer pipeline --------------------------------------------------------------------
namespace {
constexpr int kPipeStreams = 4;   // chunks of one call run on up to 4 streams
constexpr int kPipeSets = 2;      // consecutive calls alternate between two stream sets
constexpr int kPipeTickets = 4;   // calls that may be in flight per device
struct PipeTicket {
    cudaEvent_t done;
    int32_t* h_flags = nullptr;  // page-locked, kPipeStreams words
    int32_t* d_any = nullptr;    // the call's status words inside its workspace
    int used = 0;
    bool busy = false;
};
struct PipeState {  // created lazily, once per device
    std::mutex mu;  // ctypes releases the GIL: two host threads may submit / wait on one device
    cudaStream_t s[kPipeSets][kPipeStreams];
    cudaEvent_t joined[kPipeSets][kPipeStreams];
    PipeTicket ticket[kPipeTickets];
    unsigned next = 0;
    bool ready = false;
};
PipeState g_pipe[64];

int pipe_state(int device, PipeState** out) {
    if (device < 0 || device >= 64) return fail(MBQC_E_ARG, "device index %d out of range", device);
    PipeState& ps = g_pipe[device];
    std::lock_guard<std::mutex> lock(ps.mu);
    if (!ps.ready) {
        for (int g = 0; g < kPipeSets; ++g)
            for (int i = 0; i < kPipeStreams; ++i) {
                CUDA_TRY(cudaStreamCreateWithFlags(&ps.s[g][i], cudaStreamNonBlocking));
                CUDA_TRY(cudaEventCreateWithFlags(&ps.joined[g][i], cudaEventDisableTiming));
            }
        for (int t = 0; t < kPipeTickets; ++t) {
            CUDA_TRY(cudaEventCreateWithFlags(&ps.ticket[t].done, cudaEventDisableTiming));
            CUDA_TRY(cudaHostAlloc((void**)&ps.ticket[t].h_flags, kPipeStreams * sizeof(int32_t), cudaHostAllocDefault));
        }
        ps.ready = true;
    }
    *out = &ps;
    return MBQC_OK;
}
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
}  // namespace

extern "C" int64_t mbqc_host_workspace_bytes(const mbqc_plan* plan, int64_t batch, int32_t out_form) {
    if (!plan || batch < 0) return -1;
    const int k = plan->tab.n_out;
    const size_t out_elems = (out_form == MBQC_OUT_DM) ? ((size_t)1 << (2 * k)) : ((size_t)1 << k);
    const int T = plan->tab.n_angles > 0 ? plan->tab.n_angles : 1;
    return (int64_t)(align256((size_t)batch * T * sizeof(double)) + align256((size_t)batch * out_elems * sizeof(double2)) + 256);
}

extern "C" int mbqc_run_batch_sv_host_submit(const mbqc_plan* plan, const double* h_angles, int64_t angle_stride,
                                             const void* d_inputs, int32_t input_mode, int64_t batch, void* h_out,
                                             int32_t out_form, void* d_work, int64_t work_bytes, int32_t n_chunks,
                                             int32_t* ticket) {
    int rc = check_batch_args(plan, h_angles, angle_stride, d_inputs, input_mode, batch, h_out);
    if (rc) return rc;
    if (!ticket) return fail(MBQC_E_ARG, "ticket is NULL");
    if (out_form != MBQC_OUT_SV && out_form != MBQC_OUT_DM) return fail(MBQC_E_ARG, "out_form %d unknown", out_form);
    if (!d_work || work_bytes < mbqc_host_workspace_bytes(plan, batch, out_form))
        return fail(MBQC_E_ARG, "d_work too small: need %lld bytes", (long long)mbqc_host_workspace_bytes(plan, batch, out_form));
    for (int m = 0; m < plan->tab.n_steps; ++m)
        if (plan->h_steps[m].plane != MBQC_PLANE_XY || plan->h_steps[m].cond_mask)
            return fail(MBQC_E_ARG, "step %d: only the XY plane is supported on the state-vector path", m);
    PipeState* ps = nullptr;
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));
    rc = pipe_state(device, &ps);
    if (rc) return rc;
    std::lock_guard<std::mutex> pipe_lock(ps->mu);
    const unsigned slot = ps->next % kPipeTickets;
    PipeTicket& tk = ps->ticket[slot];
    if (tk.busy) return fail(MBQC_E_ARG, "%d host calls already in flight on device %d: wait for the oldest first", kPipeTickets, device);
    const int set = (int)(ps->next % kPipeSets);
    cudaStream_t* str = ps->s[set];
    *ticket = device * kPipeTickets + (int)slot;
    tk.used = 0;
    auto commit = [&]() {  // the slot is taken only once everything was queued
        ps->next++;
        tk.busy = true;
    };
    if (batch == 0) {
        CUDA_TRY(cudaEventRecord(tk.done, str[0]));
        commit();
        return MBQC_OK;
    }
    const int T = plan->tab.n_angles;
    const int Tw = T > 0 ? T : 1;
    const int k = plan->tab.n_out;
    const size_t out_elems = (out_form == MBQC_OUT_DM) ? ((size_t)1 << (2 * k)) : ((size_t)1 << k);
    char* base = (char*)d_work;
    double* d_angles = (double*)base;
    double2* d_out = (double2*)(base + align256((size_t)batch * Tw * sizeof(double)));
    int32_t* d_any = (int32_t*)((char*)d_out + align256((size_t)batch * out_elems * sizeof(double2)));
    // If the caller's output buffer is page-locked and mapped into the device address space, the
    // kernels store their (CTA-coalesced) results straight into it over PCIe and the D2H copies
    // disappear: measured 156 us vs 198 us per 65,536-sample step on B200 / PCIe 5 (profiles/).
    double2* dev_view_of_host_out = nullptr;
    static const bool direct_ok = []() { const char* e = getenv("MBQC_HOST_DIRECT"); return !(e && e[0] == '0'); }();
    if (direct_ok) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, h_out) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
            dev_view_of_host_out = (double2*)attr.devicePointer;
        else
            cudaGetLastError();  // clear the error state a plain malloc'ed pointer may leave
    }
    if (n_chunks < 1) n_chunks = 2;  // measured (profiles/): more pieces only add engine hand-offs
    if ((int64_t)n_chunks > (batch + 1023) / 1024) n_chunks = (int)((batch + 1023) / 1024);
    if (n_chunks < 1) n_chunks = 1;
    const int used = n_chunks < kPipeStreams ? n_chunks : kPipeStreams;
    for (int i = 0; i < used; ++i) CUDA_TRY(cudaMemsetAsync(d_any + i, 0, sizeof(int32_t), str[i]));
    const int64_t per = (batch + n_chunks - 1) / n_chunks;
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t lo = (int64_t)c * per;
        const int64_t hi = (lo + per < batch) ? lo + per : batch;
        if (hi <= lo) break;
        cudaStream_t st = str[c % kPipeStreams];
        if (T > 0) {
            if (angle_stride == T) {
                CUDA_TRY(cudaMemcpyAsync(d_angles + lo * T, h_angles + lo * T, (size_t)(hi - lo) * T * sizeof(double), cudaMemcpyHostToDevice, st));
            } else {
                CUDA_TRY(cudaMemcpy2DAsync(d_angles + lo * T, (size_t)T * sizeof(double), h_angles + lo * angle_stride,
                                           (size_t)angle_stride * sizeof(double), (size_t)T * sizeof(double), (size_t)(hi - lo),
                                           cudaMemcpyHostToDevice, st));
            }
        }
        SvBatchParams p;
        const void* din = d_inputs;
        if (input_mode == MBQC_INPUT_BATCH) din = (const char*)d_inputs + ((size_t)lo << plan->tab.n_in) * sizeof(double2);
        double2* dst = dev_view_of_host_out ? dev_view_of_host_out + lo * out_elems : d_out + lo * out_elems;
        fill_sv_params(p, plan, d_angles + lo * Tw, Tw, din, input_mode, hi - lo, dst, nullptr);
        p.status_any = d_any + (c % kPipeStreams);
        rc = launch_sv(p, plan, out_form, st, dev_view_of_host_out != nullptr, batch);
        if (rc) {
            // chunks already queued still read the caller's buffers: drain them before reporting
            for (int i = 0; i < kPipeStreams; ++i) cudaStreamSynchronize(str[i]);
            return rc;
        }
        if (!dev_view_of_host_out)
            CUDA_TRY(cudaMemcpyAsync((double2*)h_out + lo * out_elems, d_out + lo * out_elems, (size_t)(hi - lo) * out_elems * sizeof(double2),
                                     cudaMemcpyDeviceToHost, st));
    }
    // join: every stream's tail feeds stream 0 of the set, which fetches the status words
    for (int i = 1; i < used; ++i) {
        CUDA_TRY(cudaEventRecord(ps->joined[set][i], str[i]));
        CUDA_TRY(cudaStreamWaitEvent(str[0], ps->joined[set][i], 0));
    }
    CUDA_TRY(cudaMemcpyAsync(tk.h_flags, d_any, used * sizeof(int32_t), cudaMemcpyDeviceToHost, str[0]));
    CUDA_TRY(cudaEventRecord(tk.done, str[0]));
    tk.used = used;
    commit();
    return MBQC_OK;
}

extern "C" int mbqc_host_wait(int32_t ticket, int32_t* h_status_any) {
    if (ticket < 0 || ticket >= 64 * kPipeTickets) return fail(MBQC_E_ARG, "ticket %d out of range", ticket);
    PipeState& ps = g_pipe[ticket / kPipeTickets];
    if (!ps.ready) return fail(MBQC_E_ARG, "ticket %d: nothing was submitted on that device", ticket);
    PipeTicket& tk = ps.ticket[ticket % kPipeTickets];
    cudaEvent_t done;
    {
        std::lock_guard<std::mutex> lock(ps.mu);
        if (!tk.busy) return fail(MBQC_E_ARG, "ticket %d is not in flight", ticket);
        done = tk.done;
    }
    CUDA_TRY(cudaEventSynchronize(done));  // outside the lock: other threads keep submitting
    std::lock_guard<std::mutex> lock(ps.mu);
    tk.busy = false;
    if (h_status_any) {
        int32_t any = 0;
        for (int i = 0; i < tk.used; ++i) any |= tk.h_flags[i];
        *h_status_any = any;
    }
    return MBQC_OK;
}

extern "C" int mbqc_run_batch_sv_host(const mbqc_plan* plan, const double* h_angles, int64_t angle_stride,
                                      const void* d_inputs, int32_t input_mode, int64_t batch, void* h_out,
                                      int32_t out_form, void* d_work, int64_t work_bytes,
                                      int32_t* h_status_any, int32_t n_chunks) {
    int32_t ticket = -1;
    if (h_status_any) *h_status_any = 0;
    int rc = mbqc_run_batch_sv_host_submit(plan, h_angles, angle_stride, d_inputs, input_mode, batch, h_out, out_form,
                                           d_work, work_bytes, n_chunks, &ticket);
    if (rc) return rc;
    return mbqc_host_wait(ticket, h_status_any);
}

template <int W, bool DM>
static int launch_sv_reg_f32_w(const SvBatchParams& p, const mbqc_plan* plan, cudaStream_t st) {
    const unsigned blocks = (unsigned)((p.batch + kRegThreads - 1) / kRegThreads);
    SvRegParams rp;
    fill_reg_params(rp, p, plan);
    const size_t tables = reg_smem_tables_bytes(p.tab.n_steps, rp.reg.sign_pitch, rp.reg.n_fixed);
    const size_t tile = (size_t)kRegThreads * p.tab.n_angles * sizeof(double2);
    const int staged = (p.tab.n_angles > 0 && tables + tile <= 100 * 1024) ? 1 : 0;
    size_t smem = tables + (staged ? tile : 0);
    const size_t stage = ((size_t)kRegThreads << p.tab.n_out) * sizeof(float2);
    if (DM && stage > smem) smem = stage;
    auto kern = sv_reg_kernel_f32<W, DM>;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kRegThreads, smem, st>>>(rp, staged | (DM ? 2 : 0));
    return after_launch("sv_reg_kernel_f32");
}

extern "C" {

int mbqc_run_batch_sv_f32(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                          const void* d_inputs, int32_t input_mode, int64_t batch, void* d_out,
                          int32_t out_form, int32_t* d_status, void* stream) {
    int rc = check_batch_args(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out);
    if (rc) return rc;
    if (out_form != MBQC_OUT_SV && out_form != MBQC_OUT_DM) return fail(MBQC_E_ARG, "out_form %d unknown", out_form);
    if (plan->tab.window > MBQC_MAX_WINDOW_REG)
        return fail(MBQC_E_UNSUPPORTED, "complex64 mode covers window <= %d (got %d)", MBQC_MAX_WINDOW_REG, plan->tab.window);
    for (int m = 0; m < plan->tab.n_steps; ++m)
        if (plan->h_steps[m].plane != MBQC_PLANE_XY || plan->h_steps[m].cond_mask)
            return fail(MBQC_E_ARG, "step %d: only the XY plane is supported on the state-vector path", m);
    if (batch == 0) return MBQC_OK;
    SvBatchParams p;
    fill_sv_params(p, plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out, d_status);
    cudaStream_t st = (cudaStream_t)stream;
    const bool dm = out_form == MBQC_OUT_DM;
    switch (plan->tab.window) {
        case 1: return dm ? launch_sv_reg_f32_w<1, true>(p, plan, st) : launch_sv_reg_f32_w<1, false>(p, plan, st);
        case 2: return dm ? launch_sv_reg_f32_w<2, true>(p, plan, st) : launch_sv_reg_f32_w<2, false>(p, plan, st);
        case 3: return dm ? launch_sv_reg_f32_w<3, true>(p, plan, st) : launch_sv_reg_f32_w<3, false>(p, plan, st);
        case 4: return dm ? launch_sv_reg_f32_w<4, true>(p, plan, st) : launch_sv_reg_f32_w<4, false>(p, plan, st);
        default: return dm ? launch_sv_reg_f32_w<5, true>(p, plan, st) : launch_sv_reg_f32_w<5, false>(p, plan, st);
    }
}

// CTA size of the DM register kernel: small batches get 2-warp CTAs so that the warps spread
// evenly over the SMs (a 4096-sample C3 batch is 2048 warps = 0.86 waves of 4-warp CTAs, with
// SMs holding 3 or 4 of them; 2-warp CTAs level that to 13-14 warps per SM)
static int dm_reg_threads(int64_t batch, int n) {
    static const int forced = []() { const char* e = getenv("MBQC_DM_CTA"); return e ? atoi(e) : 0; }();
    if (forced == 32 || forced == 64 || forced == 128) return forced;
    const int64_t warps = (batch * n + 31) / 32;
    return warps < 148 * 16 * 2 ? 64 : 128;
}

static int run_batch_dm_impl(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                             const void* d_inputs, int32_t input_mode, int64_t batch, void* d_out,
                             int8_t* d_outcomes, double* d_expect, bool expect_mode, int32_t* d_status, void* stream,
                             bool z_sample = false, uint64_t z_seed = 0, uint64_t z_offset = 0) {
    int rc = check_batch_args(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out);
    if (rc) return rc;
    const int w = plan->tab.window;
    if (w > MBQC_MAX_WINDOW_SMEM_DM)
        return fail(MBQC_E_UNSUPPORTED, "batched DM covers window <= %d (got %d)", MBQC_MAX_WINDOW_SMEM_DM, w);
    bool has_z = false;
    for (int m = 0; m < plan->tab.n_steps; ++m) has_z |= plan->h_steps[m].plane == MBQC_PLANE_Z;
    if (has_z && !expect_mode && !z_sample)
        return fail(MBQC_E_UNSUPPORTED, "plane-Z steps are drawn at random by the reference outside mode='expectation' "
                                        "(np_simulator_dm.py:329-333): use mbqc_run_batch_dm_expect or mbqc_run_batch_dm_zsample");
    if (has_z && expect_mode && !d_expect) return fail(MBQC_E_ARG, "d_expect is NULL but the plan has plane-Z steps");
    if (has_z && w > MBQC_MAX_WINDOW_REG)
        return fail(MBQC_E_UNSUPPORTED, "plane-Z steps cover window <= %d (got %d)", MBQC_MAX_WINDOW_REG, w);
    if (batch == 0) return MBQC_OK;
    DmBatchParams p;
    memset(&p, 0, sizeof(p));
    p.tab = plan->tab;
    p.steps = plan->d_steps;
    p.angles = d_angles;
    p.stride = angle_stride;
    p.inputs = (const double2*)d_inputs;
    p.input_mode = input_mode;
    p.batch = batch;
    p.out = (double2*)d_out;
    p.outcomes = d_outcomes;
    p.status = d_status;
    p.expect = d_expect;
    p.z_sample = z_sample ? 1 : 0;
    p.z_seed = z_seed;
    p.z_offset = z_offset;
    if (d_expect) CUDA_TRY(cudaMemsetAsync(d_expect, 0, sizeof(double) * (size_t)batch * plan->tab.n_steps, (cudaStream_t)stream));
    // w <= 5: one lane per row of rho, registers + shuffles; w = 6: rho in shared memory
    const char* force = getenv("MBQC_DM_KERNEL");  // "smem" forces the shared-memory kernel (tests)
    if (!has_z && !force) {  // pattern-specialised kernel on the compressed state (dm_jit_src.inc)
        int jrc = MBQC_OK;
        if (mbqc_jit_dm_try_launch(plan, p, (cudaStream_t)stream, &jrc)) return jrc;
    }
    if (w <= 5 && (has_z || !(force && !strcmp(force, "smem")))) {
        const int n = 1 << w;
        const int threads = dm_reg_threads(batch, n);
        const int spb = (threads / 32) * (32 / n);
        const size_t smem = (size_t)spb * n * n * sizeof(double2);
        const unsigned blocks = (unsigned)((batch + spb - 1) / spb);
        cudaStream_t st = (cudaStream_t)stream;
        SampleParams sp;
        memset(&sp, 0, sizeof(sp));
        switch (w) {
            case 1: dm_reg_kernel<1, false><<<blocks, threads, smem, st>>>(p, sp); break;
            case 2: dm_reg_kernel<2, false><<<blocks, threads, smem, st>>>(p, sp); break;
            case 3: dm_reg_kernel<3, false><<<blocks, threads, smem, st>>>(p, sp); break;
            case 4: dm_reg_kernel<4, false><<<blocks, threads, smem, st>>>(p, sp); break;
            default:
                CUDA_TRY(cudaFuncSetAttribute(dm_reg_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                dm_reg_kernel<5, false><<<blocks, threads, smem, st>>>(p, sp);
                break;
        }
        return after_launch("dm_reg_kernel");
    }
    const unsigned ngroups = 1u << (2 * w - 2);
    int tps_log2 = ilog2_ceil((ngroups + 1) / 2);  // two 4-groups per thread
    if (tps_log2 > 8) tps_log2 = 8;
    const int tps = 1 << tps_log2;
    const size_t per_sample = ((size_t)(1u << (2 * w)) + (1u << w)) * 16;
    int spb = 256 / tps;
    while (spb > 1 && spb * per_sample > 96 * 1024) spb >>= 1;
    if (spb < 1) spb = 1;
    const size_t smem = spb * per_sample;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(dm_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)((batch + spb - 1) / spb);
    dm_smem_kernel<<<blocks, tps * spb, smem, (cudaStream_t)stream>>>(p, tps_log2, spb);
    return after_launch("dm_smem_kernel");
}

int mbqc_run_batch_dm(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                      const void* d_inputs, int32_t input_mode, int64_t batch, void* d_out,
                      int8_t* d_outcomes, int32_t* d_status, void* stream) {
    return run_batch_dm_impl(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out, d_outcomes, nullptr,
                             false, d_status, stream);
}

int mbqc_run_batch_dm_expect(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                             const void* d_inputs, int32_t input_mode, int64_t batch, void* d_out,
                             int8_t* d_outcomes, double* d_expect, int32_t* d_status, void* stream) {
    return run_batch_dm_impl(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out, d_outcomes, d_expect,
                             true, d_status, stream);
}

int mbqc_run_batch_dm_zsample(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                              const void* d_inputs, int32_t input_mode, int64_t batch, uint64_t seed,
                              uint64_t sample_offset, void* d_out, int8_t* d_outcomes, int32_t* d_status, void* stream) {
    return run_batch_dm_impl(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out, d_outcomes, nullptr,
                             false, d_status, stream, true, seed, sample_offset);
}

}  // extern "C"

namespace {
// Shared by the per-sample and the data-set entry points: p.batch samples, p.grad [batch][T].
int launch_psr_grad(const mbqc_plan* plan, const SvBatchParams& p, cudaStream_t st) {
    int rc = MBQC_OK;
    if (mbqc_jit_grad_try_launch(plan, p, st, &rc)) return rc;  // pattern-specialised kernel (sv_jit_grad_src.inc)
    const int w = plan->tab.window;
    const int64_t batch = p.batch;
    const int T = plan->tab.n_angles;
    SvRegParams rp;
    fill_reg_params(rp, p, plan);
    const size_t tables = reg_smem_tables_bytes(plan->tab.n_steps, rp.reg.sign_pitch, rp.reg.n_fixed);
    // Large batches: one thread per angle vector with prefix sharing (about half the measurements).
    // Small batches (or tiles that do not fit shared memory): one thread per (vector, parameter).
    const size_t prefix_smem = tables + (size_t)kRegThreads * ((size_t)T + ((size_t)1 << w)) * sizeof(double2);
    const char* force = getenv("MBQC_GRAD_KERNEL");  // "prefix" | "pairs" (experiments / tests)
    bool use_prefix = batch >= 32768 && prefix_smem <= 200 * 1024;
    if (force && !strcmp(force, "prefix") && prefix_smem <= 200 * 1024) use_prefix = true;
    if (force && !strcmp(force, "pairs")) use_prefix = false;
    if (use_prefix) {
        const unsigned blocks = (unsigned)((batch + kRegThreads - 1) / kRegThreads);
        auto go = [&](auto kern) -> int {
            if (prefix_smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prefix_smem));
            kern<<<blocks, kRegThreads, prefix_smem, st>>>(rp);
            return MBQC_OK;
        };
        switch (w) {
            case 1: rc = go(sv_reg_grad_prefix_kernel<1>); break;
            case 2: rc = go(sv_reg_grad_prefix_kernel<2>); break;
            case 3: rc = go(sv_reg_grad_prefix_kernel<3>); break;
            case 4: rc = go(sv_reg_grad_prefix_kernel<4>); break;
            default: rc = go(sv_reg_grad_prefix_kernel<5>); break;
        }
        if (rc) return rc;
        return after_launch("sv_reg_grad_prefix_kernel");
    }
    if (T > 128) return fail(MBQC_E_UNSUPPORTED, "fused gradient covers at most 128 angles (got %d)", T);
    const int spb = 128 / T;  // whole angle vectors per CTA
    const int threads = spb * T;
    const size_t smem = tables + (size_t)threads * sizeof(double2);
    const unsigned blocks = (unsigned)((batch + spb - 1) / spb);
    auto go = [&](auto kern) -> int {
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, threads, smem, st>>>(rp, spb);
        return MBQC_OK;
    };
    switch (w) {
        case 1: rc = go(sv_reg_grad_kernel<1>); break;
        case 2: rc = go(sv_reg_grad_kernel<2>); break;
        case 3: rc = go(sv_reg_grad_kernel<3>); break;
        case 4: rc = go(sv_reg_grad_kernel<4>); break;
        default: rc = go(sv_reg_grad_kernel<5>); break;
    }
    if (rc) return rc;
    return after_launch("sv_reg_grad_kernel");
}

int check_grad_plan(const mbqc_plan* plan, const void* d_target, double shift) {
    if (!d_target) return fail(MBQC_E_ARG, "d_target is NULL");
    if (!(shift != 0.0)) return fail(MBQC_E_ARG, "shift must be non-zero");
    const int w = plan->tab.window;
    if (w > MBQC_MAX_WINDOW_REG)
        return fail(MBQC_E_UNSUPPORTED, "fused gradient covers window <= %d (got %d)", MBQC_MAX_WINDOW_REG, w);
    for (int m = 0; m < plan->tab.n_steps; ++m)
        if (plan->h_steps[m].plane != MBQC_PLANE_XY || plan->h_steps[m].cond_mask)
            return fail(MBQC_E_ARG, "step %d: only the XY plane is supported on the state-vector path", m);
    return MBQC_OK;
}
}  // namespace

extern "C" {

// ---- sampled runs (force0 = False) -----------------------------------------------------------------
static int fill_sample_params(SampleParams& sp, const mbqc_plan* plan, int64_t batch, uint64_t seed,
                              uint64_t sample_offset, int32_t outcome_mode, int32_t correct, int8_t* d_outcomes,
                              uint32_t* d_byproducts, double* d_prob) {
    if (!plan->d_ff && plan->tab.n_steps > 0)
        return fail(MBQC_E_ARG, "the plan has no feed-forward table (mbqc_plan_set_feedforward)");
    if (outcome_mode != MBQC_OUTCOMES_SAMPLE && outcome_mode != MBQC_OUTCOMES_FORCED)
        return fail(MBQC_E_ARG, "outcome_mode %d unknown", outcome_mode);
    if (outcome_mode == MBQC_OUTCOMES_FORCED && !d_outcomes && batch > 0 && plan->tab.n_steps > 0)
        return fail(MBQC_E_ARG, "forced outcomes need d_outcomes");
    for (int m = 0; m < plan->tab.n_steps; ++m)
        if (plan->h_steps[m].plane != MBQC_PLANE_XY || plan->h_steps[m].cond_mask)
            return fail(MBQC_E_ARG, "step %d: byproduct corrections are implemented for the XY plane only", m);
    memset(&sp, 0, sizeof(sp));
    sp.ff = (const FeedForwardDev*)plan->d_ff;
    sp.seed = seed;
    sp.sample_offset = sample_offset;
    sp.outcome_mode = outcome_mode;
    sp.correct = correct;
    sp.outcomes = d_outcomes;
    sp.byproducts = d_byproducts;
    sp.prob = d_prob;
    return MBQC_OK;
}

int mbqc_run_batch_sv_sampled(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                              const void* d_inputs, int32_t input_mode, int64_t batch, uint64_t seed,
                              uint64_t sample_offset, int32_t outcome_mode, int32_t correct, void* d_out,
                              int8_t* d_outcomes, uint32_t* d_byproducts, double* d_prob, int32_t* d_status,
                              void* stream) {
    int rc = check_batch_args(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out);
    if (rc) return rc;
    const int w = plan->tab.window;
    if (w > MBQC_MAX_WINDOW_SMEM_SV)
        return fail(MBQC_E_UNSUPPORTED, "sampled runs cover window <= %d (got %d)", MBQC_MAX_WINDOW_SMEM_SV, w);
    for (int m = 0; m < plan->tab.n_steps; ++m)
        if (plan->h_steps[m].plane != MBQC_PLANE_XY || plan->h_steps[m].cond_mask)
            return fail(MBQC_E_UNSUPPORTED, "the state-vector path measures in the XY plane only (np_simulator_sv.py:54-59)");
    SampleParams sp;
    if ((rc = fill_sample_params(sp, plan, batch, seed, sample_offset, outcome_mode, correct, d_outcomes,
                                 d_byproducts, d_prob)))
        return rc;
    if (batch == 0) return MBQC_OK;
    SvBatchParams p;
    fill_sv_params(p, plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out, d_status);
    if (w > MBQC_MAX_WINDOW_REG) {  // 6..12: amplitudes in shared memory, a thread group per shot
        int tps_log2 = w - 1;
        if (tps_log2 > 8) tps_log2 = 8;
        const int tps = 1 << tps_log2;
        int spb = 256 / tps;
        if (spb < 1) spb = 1;
        const size_t smem = (size_t)spb * ((16ull << w) + 16ull * tps);
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(sv_smem_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned blocks = (unsigned)((batch + spb - 1) / spb);
        sv_smem_sample_kernel<<<blocks, tps * spb, smem, (cudaStream_t)stream>>>(p, sp, tps_log2, spb);
        return after_launch("sv_smem_sample_kernel");
    }
    SvRegParams rp;
    fill_reg_params(rp, p, plan);
    const int T = plan->tab.n_angles;
    const size_t smem = reg_smem_tables_bytes(plan->tab.n_steps, rp.reg.sign_pitch, rp.reg.n_fixed) +
                        (size_t)kRegThreads * (size_t)T * sizeof(double2);
    if (smem > 200 * 1024) return fail(MBQC_E_UNSUPPORTED, "sampled runs stage at most %d angles per pattern (got %d)", (int)((200 * 1024) / (kRegThreads * 16)), T);
    const unsigned blocks = (unsigned)((batch + kRegThreads - 1) / kRegThreads);
    cudaStream_t st = (cudaStream_t)stream;
    auto go = [&](auto kern) -> int {
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, kRegThreads, smem, st>>>(rp, sp);
        return MBQC_OK;
    };
    switch (w) {
        case 1: rc = go(sv_reg_sample_kernel<1>); break;
        case 2: rc = go(sv_reg_sample_kernel<2>); break;
        case 3: rc = go(sv_reg_sample_kernel<3>); break;
        case 4: rc = go(sv_reg_sample_kernel<4>); break;
        default: rc = go(sv_reg_sample_kernel<5>); break;
    }
    if (rc) return rc;
    return after_launch("sv_reg_sample_kernel");
}

int mbqc_run_batch_dm_sampled(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                              const void* d_inputs, int32_t input_mode, int64_t batch, uint64_t seed,
                              uint64_t sample_offset, int32_t outcome_mode, int32_t correct, void* d_out,
                              int8_t* d_outcomes, uint32_t* d_byproducts, double* d_prob, int32_t* d_status,
                              void* stream) {
    int rc = check_batch_args(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_out);
    if (rc) return rc;
    const int w = plan->tab.window;
    if (w > MBQC_MAX_WINDOW_REG)
        return fail(MBQC_E_UNSUPPORTED, "sampled runs cover window <= %d (got %d)", MBQC_MAX_WINDOW_REG, w);
    SampleParams sp;
    if ((rc = fill_sample_params(sp, plan, batch, seed, sample_offset, outcome_mode, correct, d_outcomes,
                                 d_byproducts, d_prob)))
        return rc;
    if (batch == 0) return MBQC_OK;
    DmBatchParams p;
    memset(&p, 0, sizeof(p));
    p.tab = plan->tab;
    p.steps = plan->d_steps;
    p.angles = d_angles;
    p.stride = angle_stride;
    p.inputs = (const double2*)d_inputs;
    p.input_mode = input_mode;
    p.batch = batch;
    p.out = (double2*)d_out;
    p.status = d_status;
    const int n = 1 << w;
    const int threads = dm_reg_threads(batch, n);
    const int spb = (threads / 32) * (32 / n);
    const size_t smem = (size_t)spb * n * n * sizeof(double2);
    const unsigned blocks = (unsigned)((batch + spb - 1) / spb);
    cudaStream_t st = (cudaStream_t)stream;
    switch (w) {
        case 1: dm_reg_kernel<1, true><<<blocks, threads, smem, st>>>(p, sp); break;
        case 2: dm_reg_kernel<2, true><<<blocks, threads, smem, st>>>(p, sp); break;
        case 3: dm_reg_kernel<3, true><<<blocks, threads, smem, st>>>(p, sp); break;
        case 4: dm_reg_kernel<4, true><<<blocks, threads, smem, st>>>(p, sp); break;
        default:
            CUDA_TRY(cudaFuncSetAttribute(dm_reg_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dm_reg_kernel<5, true><<<blocks, threads, smem, st>>>(p, sp);
            break;
    }
    return after_launch("dm_reg_kernel<sampled>");
}

int mbqc_psr_grad_batch(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                        const void* d_inputs, int32_t input_mode, int64_t batch,
                        const void* d_target, double shift, double* d_grad, double* d_cost,
                        int32_t* d_status, void* stream) {
    int rc = check_batch_args(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_grad);
    if (rc) return rc;
    if ((rc = check_grad_plan(plan, d_target, shift))) return rc;
    if (batch == 0 || plan->tab.n_angles == 0) return MBQC_OK;
    SvBatchParams p;
    fill_sv_params(p, plan, d_angles, angle_stride, d_inputs, input_mode, batch, nullptr, d_status);
    p.target = (const double2*)d_target;
    p.shift = shift;
    p.grad = d_grad;
    p.cost = d_cost;
    return launch_psr_grad(plan, p, (cudaStream_t)stream);
}

int mbqc_psr_grad_batch_push(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                             const void* d_inputs, int32_t input_mode, int64_t batch,
                             const void* d_target, double shift, void* const* d_results, int32_t n_results,
                             int64_t first_row, double* d_cost, int32_t* d_status, void* stream) {
    if (!d_results || n_results < 1 || n_results > 8) return fail(MBQC_E_ARG, "n_results %d not in [1, 8]", n_results);
    for (int d = 0; d < n_results; ++d)
        if (!d_results[d]) return fail(MBQC_E_ARG, "d_results[%d] is NULL", d);
    if (first_row < 0) return fail(MBQC_E_ARG, "first_row < 0");
    int rc = check_batch_args(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_results[0]);
    if (rc) return rc;
    if ((rc = check_grad_plan(plan, d_target, shift))) return rc;
    const int T = plan->tab.n_angles;
    if (batch == 0 || T == 0) return MBQC_OK;
    SvBatchParams p;
    fill_sv_params(p, plan, d_angles, angle_stride, d_inputs, input_mode, batch, nullptr, d_status);
    p.target = (const double2*)d_target;
    p.shift = shift;
    p.cost = d_cost;
    p.push_n = n_results;
    p.push_row0 = first_row;
    for (int d = 0; d < n_results; ++d) p.push_dst[d] = (double*)d_results[d];
    if (mbqc_jit_grad_try_launch(plan, p, (cudaStream_t)stream, &rc)) return rc;  // stores from the kernel itself
    // general kernels: local rows first, then one peer copy per replica (copy engines over NVLink)
    p.push_n = 0;
    p.grad = (double*)d_results[0] + first_row * T;
    if ((rc = launch_psr_grad(plan, p, (cudaStream_t)stream))) return rc;
    for (int d = 1; d < n_results; ++d)
        CUDA_TRY(cudaMemcpyAsync((double*)d_results[d] + first_row * T, p.grad, sizeof(double) * (size_t)batch * T,
                                 cudaMemcpyDefault, (cudaStream_t)stream));
    return MBQC_OK;
}

int mbqc_psr_grad_batch_multicast(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                                  const void* d_inputs, int32_t input_mode, int64_t batch,
                                  const void* d_target, double shift, void* d_result_multicast,
                                  int64_t first_row, double* d_cost, int32_t* d_status, void* stream) {
    if (!d_result_multicast) return fail(MBQC_E_ARG, "d_result_multicast is NULL");
    if (first_row < 0) return fail(MBQC_E_ARG, "first_row < 0");
    int rc = check_batch_args(plan, d_angles, angle_stride, d_inputs, input_mode, batch, d_result_multicast);
    if (rc) return rc;
    if ((rc = check_grad_plan(plan, d_target, shift))) return rc;
    if (batch == 0 || plan->tab.n_angles == 0) return MBQC_OK;
    SvBatchParams p;
    fill_sv_params(p, plan, d_angles, angle_stride, d_inputs, input_mode, batch, nullptr, d_status);
    p.target = (const double2*)d_target;
    p.shift = shift;
    p.cost = d_cost;
    p.push_n = 1;
    p.push_multicast = 1;
    p.push_row0 = first_row;
    p.push_dst[0] = (double*)d_result_multicast;
    if (mbqc_jit_grad_try_launch(plan, p, (cudaStream_t)stream, &rc)) return rc;
    // multicast addresses take multimem stores only: without the specialised kernel the caller falls back
    return fail(MBQC_E_UNSUPPORTED, "the multicast form needs the run-time specialised gradient kernel (MBQC_JIT, libnvrtc, "
                                    "every angle column read by exactly one measurement)");
}

int64_t mbqc_psr_grad_dataset_workspace_bytes(const mbqc_plan* plan, int64_t n_vectors, int64_t n_data) {
    if (!plan || n_vectors < 0 || n_data < 0) return -1;
    const int64_t n = n_vectors * n_data;
    return ((n * ((int64_t)plan->tab.n_angles * 8 + 8 + 4) + 255) / 256) * 256;
}

int mbqc_psr_grad_dataset(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                          const void* d_inputs, const void* d_targets, int64_t n_vectors, int64_t n_data,
                          double shift, double* d_grad, double* d_cost, int32_t* d_status,
                          void* d_workspace, void* stream) {
    int rc = check_batch_args(plan, d_angles, angle_stride, d_inputs, d_inputs ? MBQC_INPUT_BATCH : MBQC_INPUT_PLUS,
                              n_vectors, d_grad);
    if (rc) return rc;
    if ((rc = check_grad_plan(plan, d_targets, shift))) return rc;
    if (n_data <= 0) return fail(MBQC_E_ARG, "n_data must be positive (got %lld)", (long long)n_data);
    if (!d_workspace) return fail(MBQC_E_ARG, "d_workspace is NULL");
    const int T = plan->tab.n_angles;
    if (n_vectors == 0 || T == 0) return MBQC_OK;
    const int64_t n = n_vectors * n_data;
    double* ws_grad = (double*)d_workspace;
    double* ws_cost = ws_grad + n * T;
    int32_t* ws_status = (int32_t*)(ws_cost + n);
    SvBatchParams p;
    fill_sv_params(p, plan, d_angles, angle_stride, d_inputs, d_inputs ? MBQC_INPUT_BATCH : MBQC_INPUT_PLUS, n,
                   nullptr, ws_status);
    p.target = (const double2*)d_targets;
    p.shift = shift;
    p.grad = ws_grad;
    p.cost = ws_cost;
    p.data_count = n_data;
    if ((rc = launch_psr_grad(plan, p, (cudaStream_t)stream))) return rc;
    OptimDev none;
    memset(&none, 0, sizeof(none));
    grad_dataset_reduce_kernel<<<(unsigned)n_vectors, kReduceThreads, 0, (cudaStream_t)stream>>>(
        ws_grad, ws_cost, ws_status, n_data, T, d_grad, d_cost, d_status, 0, none);
    return after_launch("grad_dataset_reduce_kernel");
}

int mbqc_train_dataset(const mbqc_plan* plan, double* d_x, const void* d_inputs, const void* d_targets,
                       int64_t n_vectors, int64_t n_data, double shift, const mbqc_optimizer* opt,
                       int32_t first_iteration, int32_t num_iters, double* d_state, double* d_cost_history,
                       int32_t* d_status, void* d_workspace, void* stream) {
    if (!plan) return fail(MBQC_E_ARG, "plan is NULL");
    const int T = plan->tab.n_angles;
    int rc = check_batch_args(plan, d_x, T, d_inputs, d_inputs ? MBQC_INPUT_BATCH : MBQC_INPUT_PLUS, n_vectors, d_x);
    if (rc) return rc;
    if ((rc = check_grad_plan(plan, d_targets, shift))) return rc;
    if (!opt) return fail(MBQC_E_ARG, "opt is NULL");
    if (opt->kind != MBQC_OPT_ADAM && opt->kind != MBQC_OPT_SGD) return fail(MBQC_E_ARG, "optimizer kind %d unknown", opt->kind);
    if (n_data <= 0) return fail(MBQC_E_ARG, "n_data must be positive (got %lld)", (long long)n_data);
    if (num_iters < 0 || first_iteration < 0) return fail(MBQC_E_ARG, "negative iteration count");
    if (!d_state || !d_workspace) return fail(MBQC_E_ARG, "d_state / d_workspace is NULL");
    if (n_vectors == 0 || T == 0 || num_iters == 0) return MBQC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = n_vectors * n_data;
    double* ws_grad = (double*)d_workspace;
    double* ws_cost = ws_grad + n * T;
    int32_t* ws_status = (int32_t*)(ws_cost + n);
    if (d_status) CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t) * n_vectors, st));
    SvBatchParams p;
    fill_sv_params(p, plan, d_x, T, d_inputs, d_inputs ? MBQC_INPUT_BATCH : MBQC_INPUT_PLUS, n, nullptr, ws_status);
    p.target = (const double2*)d_targets;
    p.shift = shift;
    p.grad = ws_grad;
    p.cost = ws_cost;
    p.data_count = n_data;
    OptimDev o;
    memset(&o, 0, sizeof(o));
    o.kind = opt->kind;
    o.nesterov = opt->nesterov;
    o.step_size = opt->step_size;
    o.b1 = opt->b1;
    o.b2 = opt->b2;
    o.eps = opt->eps;
    o.momentum = opt->momentum;
    o.x = d_x;
    o.s0 = d_state;
    o.s1 = d_state + n_vectors * T;
    for (int it = 0; it < num_iters; ++it) {  // two launches per iteration, nothing returns to the host
        if ((rc = launch_psr_grad(plan, p, st))) return rc;
        const int t = first_iteration + it + 1;
        o.bias1 = 1.0 - std::pow(opt->b1, t);
        o.bias2 = 1.0 - std::pow(opt->b2, t);
        grad_dataset_reduce_kernel<<<(unsigned)n_vectors, kReduceThreads, 0, st>>>(
            ws_grad, ws_cost, ws_status, n_data, T, nullptr, d_cost_history ? d_cost_history + (int64_t)it * n_vectors : nullptr,
            d_status, 1, o);
        if ((rc = after_launch("grad_dataset_reduce_kernel"))) return rc;
    }
    return MBQC_OK;
}

}  // extern "C"


// ---- streaming regime --------------------------------------------------------------------------
namespace {
int stream_grid(uint64_t work_items, int threads) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t need = (work_items + threads - 1) / threads;
    const uint64_t cap = (uint64_t)sms * 8;  // persistent grid-stride CTAs, a multiple of the SM count
    return (int)(need < cap ? (need ? need : 1) : cap);
}
}  // namespace

extern "C" {

}  // extern "C"

static int make_seed(SeedDev& sd, int32_t window, int32_t n_inputs, const int32_t* input_slot,
                     const uint64_t* init_cz_mask, const void* d_input, double scale) {
    if (window < 1 || window > MBQC_MAX_WINDOW) return fail(MBQC_E_ARG, "bad window");
    if (n_inputs < 0 || n_inputs > kMaxIO || (n_inputs && !input_slot) || !init_cz_mask) return fail(MBQC_E_ARG, "bad input tables");
    memset(&sd, 0, sizeof(sd));
    sd.input = (const double2*)d_input;
    sd.n_in = n_inputs;
    sd.scale = scale;
    for (int q = 0; q < n_inputs; ++q) sd.in_slot[q] = input_slot[q];
    // cz[a] holds the slots b > a coupled to a: regroup by distance d = b - a
    for (int d = 1; d < window; ++d) {
        uint64_t m = 0;
        for (int a = 0; a + d < window; ++a)
            if ((init_cz_mask[a] >> (a + d)) & 1ull) m |= 1ull << a;
        if (m) {
            sd.dist[sd.n_dist] = d;
            sd.pair_mask[sd.n_dist] = m;
            ++sd.n_dist;
        }
    }
    return MBQC_OK;
}

extern "C" int mbqc_stream_init(void* d_state, int32_t local_bits, uint64_t index_or, int32_t window,
                     int32_t n_inputs, const int32_t* input_slot, const uint64_t* init_cz_mask,
                     const void* d_input, double scale, void* stream) {
    if (!d_state) return fail(MBQC_E_ARG, "d_state is NULL");
    if (local_bits < 0 || local_bits > window) return fail(MBQC_E_ARG, "bad window/local_bits");
    StreamInitParams p;
    int rc = make_seed(p.seed, window, n_inputs, input_slot, init_cz_mask, d_input, scale);
    if (rc) return rc;
    p.state = (double2*)d_state;
    p.n_local = 1ull << local_bits;
    p.index_or = index_or;
    stream_init_kernel<<<stream_grid(p.n_local, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return after_launch("stream_init_kernel");
}

template <bool SEED>
static int launch_stream_steps(void* d_state, const mbqc_stream_desc* desc, const SeedDev& seed, void* stream) {
    if (!d_state || !desc) return fail(MBQC_E_ARG, "NULL argument");
    if (desc->n_fused < 1 || desc->n_fused > MBQC_STREAM_MAX_FUSE) return fail(MBQC_E_ARG, "n_fused %d outside [1,%d]", desc->n_fused, MBQC_STREAM_MAX_FUSE);
    if (desc->n_ranges < 0 || desc->n_ranges > MBQC_STREAM_MAX_RANGES) return fail(MBQC_E_ARG, "n_ranges %d outside [0,%d]", desc->n_ranges, MBQC_STREAM_MAX_RANGES);
    if (desc->n_groups == 0) return MBQC_OK;
    const int grid = stream_grid(desc->n_groups, 256);
    cudaStream_t st = (cudaStream_t)stream;
    double2* s = (double2*)d_state;
    switch (desc->n_fused) {
        case 1: stream_steps_kernel<1, SEED><<<grid, 256, 0, st>>>(s, *desc, seed); break;
        case 2: stream_steps_kernel<2, SEED><<<grid, 256, 0, st>>>(s, *desc, seed); break;
        case 3: stream_steps_kernel<3, SEED><<<grid, 256, 0, st>>>(s, *desc, seed); break;
        case 4: stream_steps_kernel<4, SEED><<<grid, 256, 0, st>>>(s, *desc, seed); break;
        default: stream_steps_kernel<5, SEED><<<grid, 256, 0, st>>>(s, *desc, seed); break;
    }
    return after_launch("stream_steps_kernel");
}

extern "C" {

int mbqc_stream_steps(void* d_state, const mbqc_stream_desc* desc, void* stream) {
    SeedDev none;
    memset(&none, 0, sizeof(none));
    return launch_stream_steps<false>(d_state, desc, none, stream);
}

int mbqc_stream_steps_lanes(void* d_state, const mbqc_stream_desc* desc, void* stream) {
    if (!d_state || !desc) return fail(MBQC_E_ARG, "NULL argument");
    if (desc->n_fused < 1 || desc->n_fused > MBQC_STREAM_MAX_FUSE) return fail(MBQC_E_ARG, "n_fused %d outside [1,%d]", desc->n_fused, MBQC_STREAM_MAX_FUSE);
    if (desc->n_ranges < 0 || desc->n_ranges > MBQC_STREAM_MAX_RANGES) return fail(MBQC_E_ARG, "n_ranges %d outside [0,%d]", desc->n_ranges, MBQC_STREAM_MAX_RANGES);
    for (int j = 0; j < desc->n_fused; ++j)
        if (desc->elem_bit[j] == 0 || desc->elem_bit[j] >= 32 || (desc->elem_bit[j] & (desc->elem_bit[j] - 1)))
            return fail(MBQC_E_ARG, "lane pass: fused slot %d is not one of the 5 lane bits", j);
    for (int r = 0; r < desc->n_ranges; ++r)
        if (desc->range_pos[r] < 5) return fail(MBQC_E_ARG, "lane pass: dead slot below bit 5");
    if (desc->n_groups == 0) return MBQC_OK;
    if (desc->n_groups & 31) return fail(MBQC_E_ARG, "lane pass: element count must be a multiple of 32");
    stream_lane_kernel<<<stream_grid(desc->n_groups, 256), 256, 0, (cudaStream_t)stream>>>((double2*)d_state, *desc);
    return after_launch("stream_lane_kernel");
}

int mbqc_stream_steps_seeded(void* d_state, const mbqc_stream_desc* desc, const mbqc_stream_seed* seed,
                             void* stream) {
    if (!seed || !desc) return fail(MBQC_E_ARG, "seed/desc is NULL");
    SeedDev sd;
    int rc = make_seed(sd, seed->window, seed->n_inputs, seed->input_slot, seed->init_cz_mask, seed->d_input, seed->scale);
    if (rc) return rc;
    // symmetric initial-CZ adjacency of the fused slots, and the parity of edges inside each subset
    const int K = desc->n_fused;
    if (K < 1 || K > MBQC_STREAM_MAX_FUSE) return fail(MBQC_E_ARG, "n_fused out of range");
    int slot[MBQC_STREAM_MAX_FUSE];
    for (int j = 0; j < K; ++j) {
        const uint64_t b = desc->elem_bit[j];
        if (!b || (b & (b - 1))) return fail(MBQC_E_ARG, "elem_bit[%d] is not a single bit", j);
        slot[j] = __builtin_ctzll(b);
        uint64_t nb = (slot[j] < seed->window) ? seed->init_cz_mask[slot[j]] : 0ull;
        for (int a = 0; a < seed->window && a < slot[j]; ++a)
            if ((seed->init_cz_mask[a] >> slot[j]) & 1ull) nb |= 1ull << a;
        sd.nbr[j] = nb;
        sd.in_bit[j] = 0;
        for (int q = 0; q < seed->n_inputs; ++q)
            if (seed->input_slot[q] == slot[j]) sd.in_bit[j] = 1u << (seed->n_inputs - 1 - q);
    }
    sd.pair_parity = 0;
    for (uint32_t l = 0; l < (1u << K); ++l) {
        uint32_t par = 0;
        for (int i = 0; i < K; ++i)
            for (int j = i + 1; j < K; ++j)
                if (((l >> i) & 1u) && ((l >> j) & 1u) && ((sd.nbr[i] >> slot[j]) & 1ull)) par ^= 1u;
        sd.pair_parity |= par << l;
    }
    // the base index has the fused bits clear, so nbr[j] & base never sees another fused slot
    return launch_stream_steps<true>(d_state, desc, sd, stream);
}

int mbqc_stream_exchange(void* d_own, const void* d_peer, void* d_spare, int32_t role, double cos_t,
                         double sin_t, double scale, uint64_t nbr_mask, int32_t const_parity,
                         uint64_t n, int32_t n_ranges, const uint32_t* range_pos,
                         const uint32_t* range_width, void* stream) {
    if (n_ranges < 0 || n_ranges > MBQC_STREAM_MAX_RANGES || (n_ranges && (!range_pos || !range_width)))
        return fail(MBQC_E_ARG, "bad dead-slot ranges");
    if (!d_own || !d_peer || (role != 2 && !d_spare)) return fail(MBQC_E_ARG, "NULL buffer");
    if (role < 0 || role > 2) return fail(MBQC_E_ARG, "role %d unknown", role);
    if (n == 0) return MBQC_OK;
    ExchangeParams p;
    p.own = (double2*)d_own;
    p.peer = (const double2*)d_peer;
    p.spare = (double2*)d_spare;
    p.role = role;
    p.c = cos_t;
    p.s = sin_t;
    p.scale = scale;
    p.nbr_mask = nbr_mask;
    p.const_parity = (uint32_t)(const_parity & 1);
    p.n = n;
    p.n_ranges = n_ranges;
    for (int r = 0; r < n_ranges; ++r) {
        p.range_pos[r] = range_pos[r];
        p.range_width[r] = range_width[r];
    }
    stream_exchange_kernel<<<stream_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return after_launch("stream_exchange_kernel");
}

int mbqc_stream_gather(const void* d_state, int32_t local_bits, uint64_t index_or,
                       int32_t n_outputs, const int32_t* output_slot, void* d_out, void* stream) {
    if (!d_state || !d_out) return fail(MBQC_E_ARG, "NULL buffer");
    if (n_outputs < 0 || n_outputs > kMaxIO || (n_outputs && !output_slot)) return fail(MBQC_E_ARG, "bad output table");
    StreamGatherParams p;
    memset(&p, 0, sizeof(p));
    p.state = (const double2*)d_state;
    p.out = (double2*)d_out;
    p.index_or = index_or;
    p.local_mask = (local_bits >= 64) ? ~0ull : ((1ull << local_bits) - 1ull);
    p.n_out = n_outputs;
    for (int q = 0; q < n_outputs; ++q) p.out_slot[q] = output_slot[q];
    stream_gather_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p);
    return after_launch("stream_gather_kernel");
}

int mbqc_device_alloc(int64_t bytes, void** d_ptr) {
    if (!d_ptr || bytes < 0) return fail(MBQC_E_ARG, "bad arguments");
    *d_ptr = nullptr;
    CUDA_TRY(cudaMalloc(d_ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return MBQC_OK;
}
int mbqc_device_free(void* d_ptr) {
    if (d_ptr) CUDA_TRY(cudaFree(d_ptr));
    return MBQC_OK;
}
int mbqc_ipc_export(const void* d_ptr, void* handle64) {
    if (!d_ptr || !handle64) return fail(MBQC_E_ARG, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, (void*)d_ptr));
    return MBQC_OK;
}
int mbqc_ipc_import(const void* handle64, void** d_ptr) {
    if (!d_ptr || !handle64) return fail(MBQC_E_ARG, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MBQC_OK;
}
int mbqc_ipc_close(void* d_ptr) {
    if (d_ptr) CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
    return MBQC_OK;
}

}  // extern "C"


// ---- calculator helpers ------------------------------------------------------------------------
extern "C" {

static int split_masks(int n, const int32_t* traced, int n_traced, uint32_t* keep, uint32_t* trace) {
    if (n < 1 || n > 14) return fail(MBQC_E_ARG, "n_qubits %d outside [1,14]", n);
    uint32_t t = 0;
    for (int i = 0; i < n_traced; ++i) {
        if (traced[i] < 0 || traced[i] >= n) return fail(MBQC_E_ARG, "traced qubit %d outside the state", traced[i]);
        t |= 1u << (n - 1 - traced[i]);
    }
    *trace = t;
    *keep = ((1u << n) - 1u) & ~t;
    return MBQC_OK;
}

int mbqc_partial_trace_pure(const void* d_psi, int32_t n_qubits, const int32_t* traced, int32_t n_traced,
                            void* d_out, void* stream) {
    if (!d_psi || !d_out || (n_traced && !traced)) return fail(MBQC_E_ARG, "NULL argument");
    uint32_t keep, trace;
    int rc = split_masks(n_qubits, traced, n_traced, &keep, &trace);
    if (rc) return rc;
    trace_pure_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const double2*)d_psi, (double2*)d_out, n_qubits, keep, trace);
    return after_launch("trace_pure_kernel");
}

int mbqc_partial_trace_mixed(const void* d_rho, int32_t n_qubits, const int32_t* traced, int32_t n_traced,
                             void* d_out, void* stream) {
    if (!d_rho || !d_out || (n_traced && !traced)) return fail(MBQC_E_ARG, "NULL argument");
    uint32_t keep, trace;
    int rc = split_masks(n_qubits, traced, n_traced, &keep, &trace);
    if (rc) return rc;
    const uint64_t total = 1ull << (2 * __builtin_popcount(keep));
    trace_mixed_kernel<<<stream_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const double2*)d_rho, (double2*)d_out, n_qubits, keep, trace);
    return after_launch("trace_mixed_kernel");
}

int mbqc_pure2density(const void* d_psi, int32_t n_qubits, void* d_out, void* stream) {
    if (!d_psi || !d_out) return fail(MBQC_E_ARG, "NULL argument");
    if (n_qubits < 0 || n_qubits > 14) return fail(MBQC_E_ARG, "n_qubits %d outside [0,14]", n_qubits);
    pure2density_kernel<<<stream_grid(1ull << (2 * n_qubits), 256), 256, 0, (cudaStream_t)stream>>>((const double2*)d_psi, (double2*)d_out, n_qubits);
    return after_launch("pure2density_kernel");
}

}  // extern "C"
