// Host side of the lean register kernel (sv_lean.cuh): parameter-block prototype per plan,
// eligibility test and launch.  Separate translation unit so that kernel work rebuilds quickly.
#include <cstdlib>
#include <cstring>
#include <new>

#include "host_util.h"
#include "sv_lean.cuh"

using namespace mbqc;

static_assert(kLeanOutDirect == MBQC_LEAN_OUT_DIRECT && kLeanOutStaged == MBQC_LEAN_OUT_STAGED && kLeanOutDM == MBQC_LEAN_OUT_DM, "out modes");

// Parameter-block prototype of sv_lean_kernel: eligible patterns measure slot w-1-(m mod w) at
// step m (every reference schedule), append a qubit on a prefix of the steps and only project on
// the rest, and fit the table budget of the parameter block.
void mbqc_lean_build_proto(mbqc_plan* pl) {
    pl->lean = nullptr;
    pl->lean_fixed = 0;
    const PlanTables& t = pl->tab;
    const int w = t.window, M = t.n_steps;
    if (w < 2 || w > MBQC_MAX_WINDOW_REG || !pl->reg_periodic || M < 1 || t.n_angles < 1) return;
    const int np = 1 << (w - 1);
    if (M > kLeanMaxSteps || M * np > kLeanMaxSignWords) return;
    int n_full = 0;
    while (n_full < M && (pl->h_steps[n_full].flags & MBQC_STEP_APPEND)) ++n_full;
    for (int m = n_full; m < M; ++m)
        if (pl->h_steps[m].flags & MBQC_STEP_APPEND) return;
    if (M - n_full > w) return;
    LeanParams* lp = new (std::nothrow) LeanParams();
    if (!lp) return;
    memset(lp, 0, sizeof(*lp));
    lp->fixed = pl->d_reg_fixed;
    lp->n_angles = t.n_angles;
    lp->n_steps = M;
    lp->n_out = t.n_out;
    lp->n_in = t.n_in;
    lp->n_full = n_full;
    lp->init_sign = t.init_sign;
    lp->init_scale = t.init_scale;
    memcpy(lp->init_src, t.init_src, sizeof(lp->init_src));
    memcpy(lp->out_dst, t.out_dst, sizeof(lp->out_dst));
    int n_fixed = 0;
    for (int m = 0; m < M; ++m) {
        const StepDev& d = pl->h_steps[m];
        if (d.angle_idx >= 0) lp->colofs[m] = (uint32_t)d.angle_idx * 8u;
        else lp->colofs[m] = kLeanFixedBit | (uint32_t)(n_fixed++);  // same order as d_reg_fixed
        int pidx = 0;
        for (uint32_t i = 0; i < (1u << w); ++i) {
            if ((i >> d.slot) & 1u) continue;
            const uint32_t j = i | (1u << d.slot);
            lp->signs[(size_t)m * np + pidx] = ((d.flipmask >> j) & 1u) ? 0x80000000u : 0u;
            ++pidx;
        }
    }
    pl->lean = lp;
    pl->lean_fixed = n_fixed > 0;
}


void mbqc_lean_free_proto(mbqc_plan* plan) {
    delete plan->lean;
    plan->lean = nullptr;
}

// ---- lean register kernel (sv_lean.cuh) ----
template <int W, int CTA, bool FIXED, int OUT>
static int launch_lean_inst(const LeanParams& lp, size_t smem, cudaStream_t st) {
    auto kern = sv_lean_kernel<W, CTA, FIXED, OUT>;
    if (smem > 40 * 1024) {  // static shared memory (trig table, barriers) counts against the 48 KB default too
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return mbqc_cuda_error(e, "cudaFuncSetAttribute(sv_lean_kernel)");
    }
    const unsigned blocks = (unsigned)((lp.batch + CTA - 1) / CTA);
    kern<<<blocks, CTA, smem, st>>>(lp);
    return mbqc_after_launch("sv_lean_kernel");
}

template <int W>
static int launch_lean_w(const LeanParams& lp, bool fixed, int out_mode, int cta, size_t smem, cudaStream_t st) {
    if (cta == 64 && !fixed && out_mode == kLeanOutDirect) return launch_lean_inst<W, 64, false, kLeanOutDirect>(lp, smem, st);
    if (fixed) {
        switch (out_mode) {
            case kLeanOutDirect: return launch_lean_inst<W, 128, true, kLeanOutDirect>(lp, smem, st);
            case kLeanOutStaged: return launch_lean_inst<W, 128, true, kLeanOutStaged>(lp, smem, st);
            default: return launch_lean_inst<W, 128, true, kLeanOutDM>(lp, smem, st);
        }
    }
    switch (out_mode) {
        case kLeanOutDirect: return launch_lean_inst<W, 128, false, kLeanOutDirect>(lp, smem, st);
        case kLeanOutStaged: return launch_lean_inst<W, 128, false, kLeanOutStaged>(lp, smem, st);
        default: return launch_lean_inst<W, 128, false, kLeanOutDM>(lp, smem, st);
    }
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// 0 = not eligible (caller falls back to sv_reg_kernel), 1 = launched (rc holds the result)
int mbqc_lean_try_launch(const SvBatchParams& p, const mbqc_plan* plan, int out_mode, cudaStream_t st, int* rc) {
    static const int disabled = env_int("MBQC_SV_KERNEL_REG", 0);  // 1: always the general register kernel
    static const int cta_env = env_int("MBQC_LEAN_CTA", 0);
    if (disabled || !plan->lean) return 0;
    const int T = p.tab.n_angles;
    if (p.stride != T || ((uintptr_t)p.angles & 15u)) return 0;
    int cta = 128;
    if (!plan->lean_fixed && out_mode == kLeanOutDirect) {
        // MBQC_LEAN_CTA=64: 64-thread CTAs (measured equal to 128 on B200 at 65,536 .. 4M samples)
        if (cta_env == 64) cta = 64;
    }
    size_t smem = (size_t)cta * T * sizeof(double);
    if (out_mode != kLeanOutDirect) {
        const size_t stage = ((size_t)cta << p.tab.n_out) * sizeof(double2);
        if (stage > smem) smem = stage;
    }
    if (smem > 96 * 1024) return 0;
    LeanParams lp = *plan->lean;
    lp.angles = p.angles;
    lp.out = p.out;
    lp.status = p.status;
    lp.status_any = p.status_any;
    lp.inputs = p.inputs;
    lp.batch = p.batch;
    lp.input_mode = p.input_mode;
    switch (p.tab.window) {
        case 2: *rc = launch_lean_w<2>(lp, plan->lean_fixed, out_mode, cta, smem, st); break;
        case 3: *rc = launch_lean_w<3>(lp, plan->lean_fixed, out_mode, cta, smem, st); break;
        case 4: *rc = launch_lean_w<4>(lp, plan->lean_fixed, out_mode, cta, smem, st); break;
        default: *rc = launch_lean_w<5>(lp, plan->lean_fixed, out_mode, cta, smem, st); break;
    }
    return 1;
}

