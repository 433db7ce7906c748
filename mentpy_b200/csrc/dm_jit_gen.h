// Host-side generator of the preamble of the specialised density-matrix kernel (dm_jit_src.inc):
// the register / lane layout of every step, the slot exchanges, CZ sign tables, angle columns,
// planes, channel constants and output tables of one plan, as C++ definitions.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "dm_params.cuh"

namespace mbqc {

struct DmJitShape {
    int lb = 0, nlb = 0;      // register / lane slots
    int cta = 64, minblocks = 1;
    int xbufs = 2, ost_ofs = 0, ost_extra = 0;  // exchange buffers per warp; place of the output block (double2 units)
    size_t smem = 0;          // dynamic shared memory of one CTA
    int lanes() const { return 1 << (2 * nlb); }
    int samples_per_cta() const { return (cta / 32) * (32 / lanes()); }
};

constexpr int kDmJitPitch = 33;  // double2 units: destination-major exchange buffer, conflict-free reads

inline void dm_appendf(std::string& s, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    s += buf;
}

// Patterns the specialised kernel covers: the reference schedule (slot w-1-(m mod w) at step m),
// window 2..5, projective planes only (plane Z / expectation mode stays on dm_reg_kernel).
inline bool dm_jit_shape(const mbqc_plan* plan, DmJitShape& sh, int lb_request = 0) {
    const PlanTables& t = plan->tab;
    const int w = t.window, M = t.n_steps;
    if (w < 2 || w > 5 || M < 1 || !plan->reg_periodic) return false;
    for (int m = 0; m < M; ++m)
        if (plan->h_steps[m].plane == MBQC_PLANE_Z || plan->h_steps[m].cond_mask) return false;
    // lanes per sample = 4^(w-1-lb) <= 32; registers per lane = 2 * 4^lb doubles
    sh.lb = (w == 5) ? 2 : 1;
    if (lb_request >= 1 && lb_request <= 2 && w - 1 - lb_request >= 0 && w - 1 - lb_request <= 2) sh.lb = lb_request;
    sh.nlb = w - 1 - sh.lb;
    sh.cta = 64;
    // per warp: exchange buffer(s) (the seed state and the final sigma share buffer 0) + the gathered
    // output block when a channel has to run over it (in buffer 1 where there is one and it fits)
    const size_t nreg = (size_t)1 << (2 * sh.lb), xbuf = nreg * kDmJitPitch;
    const size_t warps = sh.cta / 32, spw = 32 / sh.lanes(), spb = sh.samples_per_cta();
    sh.xbufs = sh.lb == 1 ? 2 : 1;
    const size_t ost = (t.has_noise && t.n_out > 0) ? spw << (2 * t.n_out) : 0;
    if (ost == 0) {
        sh.ost_ofs = sh.ost_extra = 0;
    } else if (sh.xbufs == 2 && ost <= xbuf) {
        sh.ost_ofs = (int)xbuf;
        sh.ost_extra = 0;
    } else {
        sh.ost_ofs = (int)(sh.xbufs * xbuf);
        sh.ost_extra = (int)ost;
    }
    if ((spw << w) > xbuf || (spw << (2 * (w - 1))) > xbuf) return false;
    sh.smem = 16 * (warps * (sh.xbufs * xbuf + sh.ost_extra) + spb * (size_t)M * 2);
    if (sh.smem > 160 * 1024) return false;
    // registers: 2 * 4^lb doubles of state plus the working set of one group
    sh.minblocks = sh.lb == 1 ? 8 : 6;  // measured: 16 (64 registers) costs 2-7 % at every batch size
    if (const char* e = getenv("MBQC_DM_MINBLOCKS")) sh.minblocks = atoi(e) > 0 ? atoi(e) : sh.minblocks;  // kernel work
    return true;
}

inline std::string dm_jit_preamble(const mbqc_plan* plan, const DmJitShape& sh) {
    const PlanTables& t = plan->tab;
    const int w = t.window, M = t.n_steps, LB = sh.lb, NLB = sh.nlb, NP = w - 1;
    auto slot_of = [&](int m) { return w - 1 - (m % w); };
    auto next_use = [&](int slot, int after) {
        for (int m = after + 1; m < M; ++m)
            if (slot_of(m) == slot) return m;
        return 1 << 30;
    };
    // layout produced by step 0 (= consumed by step 1): the slots measured soonest are register slots
    std::vector<int> pos_slot;
    for (int s = 0; s < w; ++s)
        if (s != slot_of(0)) pos_slot.push_back(s);
    for (int a = 0; a < NP; ++a)
        for (int b = a + 1; b < NP; ++b)
            if (next_use(pos_slot[b], 0) < next_use(pos_slot[a], 0)) std::swap(pos_slot[a], pos_slot[b]);
    const std::vector<int> pos_slot1 = pos_slot;
    std::vector<int> measpos(M, 0), e(M, 0), xchg(M, 0), xreg(M, 0), xlane(M, 0), xbuf(M, 0);
    std::vector<unsigned> sgnreg(M, 0u), sgnlane(M, 0u);
    int n_xchg = 0;
    for (int m = 1; m < M; ++m) {
        const int sm = slot_of(m);
        int p = 0;
        while (pos_slot[p] != sm) ++p;
        if (p >= LB) {  // bring the measured slot into the registers: evict the slot needed last
            int k = 0;
            for (int c = 1; c < LB; ++c)
                if (next_use(pos_slot[c], m) > next_use(pos_slot[k], m)) k = c;
            std::swap(pos_slot[k], pos_slot[p]);
            xchg[m] = 1;
            xreg[m] = k;
            xlane[m] = p - LB;
            xbuf[m] = n_xchg++ & 1;
            p = k;
        }
        measpos[m] = p;
        const uint64_t mask = plan->h_steps[m - 1].nbr_mask;  // 0 for steps without an append
        e[m] = (int)((mask >> sm) & 1ull);
        for (int k = 0; k < LB; ++k)
            if (k != p && ((mask >> pos_slot[k]) & 1ull)) sgnreg[m] |= 1u << k;
        for (int j = 0; j < NLB; ++j)
            if ((mask >> pos_slot[LB + j]) & 1ull) sgnlane[m] |= 1u << j;
        pos_slot[p] = slot_of(m - 1);  // the qubit appended by the previous step becomes explicit
    }
    // canonical (slot-ordered) bit of every position in the final layout
    const int s_last = slot_of(M - 1);
    std::vector<int> finalbit(NP, 0);
    for (int p = 0; p < NP; ++p)
        for (int q = 0; q < NP; ++q)
            if (pos_slot[q] < pos_slot[p]) ++finalbit[p];
    auto compress = [&](uint64_t idx) {  // drop bit s_last
        return (int)(((idx >> (s_last + 1)) << s_last) | (idx & ((1ull << s_last) - 1ull)));
    };
    const uint64_t mask_last = plan->h_steps[M - 1].nbr_mask;

    std::string s;
    dm_appendf(s, "#define JW %d\n#define JM %d\n#define JLB %d\n#define JNLB %d\n#define JCTA %d\n#define JMINBLOCKS %d\n", w, M, LB,
               NLB, sh.cta, sh.minblocks);
    dm_appendf(s, "#define JNOUT %d\n#define JNIN %d\n#define JPITCH %d\n#define JNOISE %d\n", t.n_out, t.n_in, kDmJitPitch,
               t.has_noise ? 1 : 0);
    dm_appendf(s, "#define JXBUFS %d\n#define JOST_OFS %d\n#define JOST_EXTRA %d\n", sh.xbufs, sh.ost_ofs, sh.ost_extra);
    const mbqc_noise& nz = t.noise;
    dm_appendf(s, "#define JPOP0 %a\n#define JPOP1 %a\n#define JPOP2 %a\n#define JPOP3 %a\n#define JCOHG %a\n#define JCOHD %a\n",
               t.has_noise ? nz.pop[0] : 1.0, t.has_noise ? nz.pop[1] : 0.0, t.has_noise ? nz.pop[2] : 0.0,
               t.has_noise ? nz.pop[3] : 1.0, t.has_noise ? nz.coh_g : 1.0, t.has_noise ? nz.coh_d : 0.0);
    dm_appendf(s, "#define JINIT_SCALE %a\n#define JPLUS_AMP %a\n", t.init_scale, t.plus_amp);
    dm_appendf(s, "constexpr int kS0 = %d;\nconstexpr unsigned kInitSign = %uu;\n", slot_of(0), t.init_sign);
    auto int_table = [&](const char* decl, const std::vector<int>& v) {
        s += decl;
        s += " = {";
        for (int x : v) dm_appendf(s, "%d,", x);
        s += "};\n";
    };
    auto uint_table = [&](const char* decl, const std::vector<unsigned>& v) {
        s += decl;
        s += " = {";
        for (unsigned x : v) dm_appendf(s, "%uu,", x);
        s += "};\n";
    };
    int_table("constexpr int kMeasPos[JM]", measpos);
    int_table("constexpr int kE[JM]", e);
    uint_table("constexpr unsigned kSgnReg[JM]", sgnreg);
    uint_table("constexpr unsigned kSgnLane[JM]", sgnlane);
    int_table("constexpr int kXchg[JM]", xchg);
    int_table("constexpr int kXchgReg[JM]", xreg);
    int_table("constexpr int kXchgLane[JM]", xlane);
    int_table("constexpr int kXchgBuf[JM]", xbuf);
    int_table("constexpr int kPosSlot1[]", pos_slot1);
    int_table("constexpr int kFinalBit[]", finalbit);
    std::vector<int> aidx(M), plane(M), isrc(1 << w), orow(1 << t.n_out), oneg(1 << t.n_out);
    std::string fc = "__device__ const double kFixedCosRt[JM] = {", fs = "__device__ const double kFixedSinRt[JM] = {";
    std::string fz = "__device__ const double kFixedZRt[JM] = {";
    for (int m = 0; m < M; ++m) {
        const StepDev& d = plan->h_steps[m];
        aidx[m] = d.angle_idx;
        plane[m] = d.plane;
        dm_appendf(fc, "%a,", d.angle_idx >= 0 ? 1.0 : d.fc);
        dm_appendf(fs, "%a,", d.angle_idx >= 0 ? 0.0 : d.fs);
        dm_appendf(fz, "%a,", d.fz);
    }
    s += fc + "};\n" + fs + "};\n" + fz + "};\n";
    for (int i = 0; i < (1 << w); ++i) isrc[i] = t.init_src[i];
    for (int d = 0; d < (1 << t.n_out); ++d) {
        const uint64_t ri = output_state_index(t, (uint32_t)d);
        orow[d] = compress(ri);
        oneg[d] = (int)(((ri >> s_last) & 1ull) & parity64(ri & mask_last));
    }
    int_table("__device__ const int kAngleIdxRt[JM]", aidx);
    int_table("__device__ const int kPlaneRt[JM]", plane);
    int_table("__device__ const int kInitSrcRt[]", isrc);
    int_table("__device__ const int kOutRowRt[]", orow);
    int_table("__device__ const int kOutNegRt[]", oneg);
    std::string spec = "#define JSTEPS_SPEC", exact = "#define JSTEPS_EXACT";
    for (int m = 0; m < M; ++m) {
        dm_appendf(spec, " dm_step_spec<%d>(vr, vi, L, trc, rare);", m);
        dm_appendf(exact, " dm_step_exact<%d>(er, ei, LE, trc, took1, bad, outc);", m);
    }
    s += spec + "\n" + exact + "\n";
    return s;
}

}  // namespace mbqc
