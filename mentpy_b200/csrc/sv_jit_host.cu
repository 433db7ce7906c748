// Run-time specialisation of the batched state-vector kernel (sv_jit_src.inc) for one plan:
// source generation, NVRTC compilation (libnvrtc is loaded with dlopen -- the library has no link
// dependency on it), cubin cache (process-wide map + files), launch through the runtime's
// library API.  When NVRTC is not available or a compilation fails the caller simply continues
// with the ahead-of-time kernels (sv_lean_kernel / sv_reg_kernel): both are CUDA paths.
//
// Switches (read once): MBQC_JIT=0 disables, MBQC_JIT=force specialises every eligible call,
// default: calls with batch >= MBQC_JIT_MIN_BATCH (16384).  MBQC_JIT_CACHE names the directory of
// the cubin cache (default $HOME/.cache/mentpy_b200; "off" disables the files).
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "dm_jit_gen.h"
#include "host_util.h"
#include "sv_lean.cuh"

using namespace mbqc;

namespace {

const char* kJitSource =
#include "sv_jit_src.inc"
    ;
const char* kJitGradSource =
#include "sv_jit_grad_src.inc"
    ;
const char* kJitDmSource =
#include "dm_jit_src.inc"
    ;
constexpr int kKindRun = 0, kKindGrad = 1, kKindDm = 2;
const char* kind_source(int kind) { return kind == kKindGrad ? kJitGradSource : (kind == kKindDm ? kJitDmSource : kJitSource); }
const char* kind_entry(int kind) { return kind == kKindGrad ? "mbqc_jit_grad" : (kind == kKindDm ? "mbqc_jit_dm" : "mbqc_jit_sv"); }
const char* kind_tag(int kind) { return kind == kKindGrad ? "grad" : (kind == kKindDm ? "dm" : "sv"); }

// ---- NVRTC through dlopen --------------------------------------------------------------------
struct Nvrtc {
    void* handle = nullptr;
    std::string where, error;
    int major = 0, minor = 0;
    decltype(&nvrtcCreateProgram) createProgram = nullptr;
    decltype(&nvrtcCompileProgram) compileProgram = nullptr;
    decltype(&nvrtcDestroyProgram) destroyProgram = nullptr;
    decltype(&nvrtcGetCUBINSize) getCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) getCUBIN = nullptr;
    decltype(&nvrtcGetProgramLogSize) getLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) getLog = nullptr;
    decltype(&nvrtcVersion) version = nullptr;
    bool ok() const { return handle != nullptr; }
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> names;
        if (const char* e = getenv("MBQC_NVRTC_PATH")) names.push_back(e);
        for (const char* s : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                              "/usr/local/cuda/lib64/libnvrtc.so"})
            names.push_back(s);
        for (const auto& nm : names) {
            n.handle = dlopen(nm.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (n.handle) {
                n.where = nm;
                break;
            }
        }
        if (!n.handle) {
            n.error = "libnvrtc not found (set MBQC_NVRTC_PATH)";
            return;
        }
#define MBQC_SYM(field, name)                                             \
    n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.handle, name)); \
    if (!n.field) {                                                       \
        n.error = std::string("libnvrtc lacks ") + name;                  \
        dlclose(n.handle);                                                \
        n.handle = nullptr;                                               \
        return;                                                           \
    }
        MBQC_SYM(createProgram, "nvrtcCreateProgram")
        MBQC_SYM(compileProgram, "nvrtcCompileProgram")
        MBQC_SYM(destroyProgram, "nvrtcDestroyProgram")
        MBQC_SYM(getCUBINSize, "nvrtcGetCUBINSize")
        MBQC_SYM(getCUBIN, "nvrtcGetCUBIN")
        MBQC_SYM(getLogSize, "nvrtcGetProgramLogSize")
        MBQC_SYM(getLog, "nvrtcGetProgramLog")
        MBQC_SYM(version, "nvrtcVersion")
#undef MBQC_SYM
        n.version(&n.major, &n.minor);
    });
    return n;
}

// ---- source generation ---------------------------------------------------------------------------
struct Variant {
    int out_mode, cta;
    int kind = kKindRun;  // kKindRun: mbqc_jit_sv, kKindGrad: mbqc_jit_grad
    bool operator<(const Variant& o) const {
        if (kind != o.kind) return kind < o.kind;
        return out_mode != o.out_mode ? out_mode < o.out_mode : cta < o.cta;
    }
};

void appendf(std::string& s, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    s += buf;
}

// everything the kernel needs to know about the plan, as preprocessor / constexpr definitions
std::string make_preamble(const mbqc_plan* plan, Variant v) {
    if (v.kind == kKindDm) {
        DmJitShape sh;
        if (!dm_jit_shape(plan, sh, v.out_mode)) return std::string();
        return dm_jit_preamble(plan, sh);
    }
    const LeanParams& lp = *plan->lean;
    const int w = plan->tab.window, M = lp.n_steps, np = 1 << (w - 1), n = 1 << w;
    int minblocks = (w <= 3) ? 1024 / v.cta : (w == 4 ? 640 / v.cta : 384 / v.cta);
    // gradient kernel: two working buffers; at w = 5 its shared memory (96 KB per 128 threads) admits 256
    // threads per SM anyway, so the register cap is lifted to match (1.82 -> 1.69 ms on C4)
    if (v.kind == kKindGrad) minblocks = (w <= 3) ? 1024 / v.cta : (w == 4 ? 640 / v.cta : 256 / v.cta);
    if (const char* e = getenv(v.kind == kKindGrad ? "MBQC_GRAD_MINBLOCKS" : "MBQC_SV_MINBLOCKS"))  // kernel work
        minblocks = atoi(e) > 0 ? atoi(e) : minblocks;
    std::string s;
    appendf(s, "#define JW %d\n#define JM %d\n#define JNFULL %d\n#define JT %d\n#define JNOUT %d\n#define JNIN %d\n", w, M,
            lp.n_full, lp.n_angles, lp.n_out, lp.n_in);
    appendf(s, "#define JCTA %d\n#define JMINBLOCKS %d\n#define JOUT %d\n#define JPHASE %s\n", v.cta, minblocks, v.out_mode,
            v.out_mode == MBQC_LEAN_OUT_DM ? "false" : "true");
    if (v.kind == kKindGrad) appendf(s, "#define JPUSH %d\n", v.out_mode);  // 1..3: replicated result (stores | bulk copies, full wait | read wait)
    appendf(s, "#define JTRIG_INV %a\n#define JTRIG_C1 %a\n#define JTRIG_C2 %a\n", (double)MBQC_TRIG128_INV, (double)MBQC_TRIG128_C1,
            (double)MBQC_TRIG128_C2);
    std::string col = "constexpr unsigned kColOfs[JM] = {", sgn = "constexpr unsigned kSignMask[JM] = {",
                fix = "constexpr int kFixedIdx[JM] = {", fc = "constexpr double kFixedCos[] = {", fs = "constexpr double kFixedSin[] = {";
    int n_fixed = 0;
    for (int m = 0; m < M; ++m) {
        const bool fixed = (lp.colofs[m] & kLeanFixedBit) != 0;
        appendf(col, "%uu,", fixed ? 0u : lp.colofs[m]);
        unsigned mask = 0;
        for (int p = 0; p < np; ++p)
            if (lp.signs[(size_t)m * np + p]) mask |= 1u << p;
        appendf(sgn, "%uu,", mask);
        appendf(fix, "%d,", fixed ? n_fixed : -1);
        if (fixed) {
            appendf(fc, "%a,", plan->h_steps[m].fc);
            appendf(fs, "%a,", plan->h_steps[m].fs);
            ++n_fixed;
        }
    }
    if (n_fixed == 0) {
        fc += "0.0";
        fs += "0.0";
    }
    s += col + "};\n" + sgn + "};\n" + fix + "};\n" + fc + "};\n" + fs + "};\n";
    appendf(s, "constexpr unsigned kInitSign = %uu;\n", lp.init_sign);
    std::string src = "constexpr int kInitSrc[] = {", dst = "constexpr int kOutDst[] = {";
    for (int i = 0; i < n; ++i) {
        appendf(src, "%d,", (int)lp.init_src[i]);
        appendf(dst, "%d,", (int)lp.out_dst[i]);
    }
    s += src + "};\n" + dst + "};\n";
    if (v.kind == kKindRun) {
        std::string steps = "#define JSTEPS";
        for (int m = 0; m < M; ++m) appendf(steps, " jit_renorm<%d>(re, im, zr, zi); jit_step<%d>(row, s_trig, re, im, zr, zi, big);", m, m);
        s += steps + "\n";
    } else {
        // run-time indexable copies of two tables + the switch bodies of the gradient kernel
        std::string fx = "__constant__ int kFixedIdxRt[JM] = {", cl = "__constant__ int kColRt[JM] = {";
        // measurement m reads the buffer m & 1 (E = even, O = odd) and writes the other one; the
        // power-of-two damping every 8 measurements is part of the case body.  STEP: working state;
        // YSTEP: prefix (shared memory) -> u-part in the working state; PSTEP: prefix -> prefix.
        std::string cs = "#define JCASES_STEP", cy = "#define JCASES_YSTEP", cp = "#define JCASES_PSTEP";
        int nf = 0;
        for (int m = 0; m < M; ++m) {
            const bool fixed = (lp.colofs[m] & kLeanFixedBit) != 0;
            appendf(fx, "%d,", fixed ? nf++ : -1);
            appendf(cl, "%d,", fixed ? 0 : (int)(lp.colofs[m] / 8));
            if (m == 0) continue;  // measurement 0 reads the full seed state: handled outside the switches
            const bool odd = m & 1;
            const char* src = odd ? "wOr, wOi" : "wEr, wEi";
            const char* dst = odd ? "wEr, wEi" : "wOr, wOi";
            std::string damp = ((m & 7) == 7) ? std::string(" damp(") + dst + ");" : std::string();
            appendf(cs, " case %d: step_angle<%d>(cs, c, s); jit_cstep<%d, false>(%s, %s, c, s);%s break;", m, m, m, src, dst, damp.c_str());
            if (!fixed)
                appendf(cy, " case %d: load_prefix(pfx, %s); step_angle<%d>(cs, c, s); jit_cstep<%d, true>(%s, %s, c, s);%s break;", m, src,
                        m, m, src, dst, damp.c_str());
            appendf(cp, " case %d: load_prefix(pfx, %s); step_angle<%d>(cs, c, s); jit_cstep<%d, false>(%s, %s, c, s);%s store_prefix(pfx, %s); break;",
                    m, src, m, m, src, dst, damp.c_str(), dst);
        }
        s += cp + "\n";
        s += fx + "};\n" + cl + "};\n" + cs + "\n" + cy + "\n";
        // compressed-state bookkeeping: which register-index bit holds the slot measured at step m,
        // and the pending CZ signs of step m-1 seen from the register pairs of step m
        std::vector<int> pos2slot;
        const int s0 = w - 1;
        for (int sl = 0; sl < w; ++sl)
            if (sl != s0) pos2slot.push_back(sl);
        auto full_index = [&](int r) {
            int i = 0;
            for (int q = 0; q < w - 1; ++q)
                if ((r >> q) & 1) i |= 1 << pos2slot[q];
            return i;
        };
        auto pending_mask = [&](int m) -> uint64_t {  // CZ neighbours of the qubit appended at step m
            return (plan->h_steps[m].flags & MBQC_STEP_APPEND) ? plan->h_steps[m].nbr_mask : 0ull;
        };
        std::string qb = "constexpr int kQBit[JM] = {0,", sa = "constexpr unsigned kSA[JM] = {0u,", sb = "constexpr unsigned kSB[JM] = {0u,";
        const int groups = (1 << (w - 1)) / 2;
        for (int m = 1; m < M; ++m) {
            const int sm = w - 1 - (m % w), sprev = w - 1 - ((m - 1) % w);
            int q = -1;
            for (int t = 0; t < w - 1; ++t)
                if (pos2slot[t] == sm) q = t;
            const uint64_t mask = pending_mask(m - 1);
            unsigned ma = 0, mb = 0;
            for (int g = 0; g < groups; ++g) {
                const int r0 = ((g >> q) << (q + 1)) | (g & ((1 << q) - 1)), r1 = r0 | (1 << q);
                if (__builtin_parityll((uint64_t)full_index(r0) & mask)) ma |= 1u << g;
                if (__builtin_parityll((uint64_t)full_index(r1) & mask)) mb |= 1u << g;
            }
            appendf(qb, "%d,", q);
            appendf(sa, "%uu,", ma);
            appendf(sb, "%uu,", mb);
            pos2slot[q] = sprev;
        }
        s += qb + "};\n" + sa + "};\n" + sb + "};\n";
        // outputs: register and pending sign of output index d in the final compressed state
        const int slast = w - 1 - ((M - 1) % w);
        const uint64_t mlast = pending_mask(M - 1);
        std::vector<int> oreg(1 << lp.n_out, 0), oneg(1 << lp.n_out, 0);
        for (int i = 0; i < n; ++i) {
            const int d = lp.out_dst[i];
            if (d < 0) continue;
            int r = 0;
            for (int t = 0; t < w - 1; ++t)
                if ((i >> pos2slot[t]) & 1) r |= 1 << t;
            oreg[d] = r;
            oneg[d] = ((i >> slast) & 1) ? __builtin_parityll((uint64_t)(i & ~(1 << slast)) & mlast) : 0;
        }
        std::string orr = "constexpr int kOutReg[] = {", onn = "constexpr bool kOutNeg[] = {";
        for (size_t d = 0; d < oreg.size(); ++d) {
            appendf(orr, "%d,", oreg[d]);
            appendf(onn, "%s,", oneg[d] ? "true" : "false");
        }
        s += orr + "};\n" + onn + "};\n";
    }
    return s;
}

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char c : s) {
        h ^= c;
        h *= 1099511628211ull;
    }
    return h;
}

// ---- cache ---------------------------------------------------------------------------------------
struct Loaded {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    bool failed = false;
};

struct JitState {
    std::mutex mu;
    std::map<std::pair<uint64_t, int>, Loaded> kernels;  // (source hash, device) -> loaded kernel
    long long compiled = 0, from_disk = 0, failures = 0;
    std::string last_error;
};
JitState& state() {
    static JitState s;
    return s;
}

std::string cache_dir() {
    const char* e = getenv("MBQC_JIT_CACHE");
    if (e && !strcmp(e, "off")) return "";
    std::string d;
    if (e && *e) d = e;
    else if (const char* h = getenv("HOME")) d = std::string(h) + "/.cache/mentpy_b200";
    else return "";
    mkdir(d.substr(0, d.find_last_of('/')).c_str(), 0755);
    mkdir(d.c_str(), 0755);
    return d;
}

bool read_file(const std::string& path, std::vector<char>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? n : 0);
    const bool ok = n > 0 && fread(out.data(), 1, n, f) == (size_t)n;
    fclose(f);
    return ok;
}

void write_file_atomic(const std::string& path, const std::vector<char>& data) {
    const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    if (ok) rename(tmp.c_str(), path.c_str());
    else unlink(tmp.c_str());
}

// compile `source` for sm_100a; returns false and fills `err` on failure
bool compile_cubin(const std::string& source, std::vector<char>& cubin, std::string& err) {
    Nvrtc& n = nvrtc();
    if (!n.ok()) {
        err = n.error;
        return false;
    }
    nvrtcProgram prog = nullptr;
    if (n.createProgram(&prog, source.c_str(), "mbqc_jit_sv.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
        err = "nvrtcCreateProgram failed";
        return false;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--extra-device-vectorization"};
    const nvrtcResult rc = n.compileProgram(prog, 3, opts);
    if (rc != NVRTC_SUCCESS) {
        size_t ls = 0;
        n.getLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.getLog(prog, &log[0]);
        err = "nvrtcCompileProgram failed: " + log.substr(0, 1500);
        n.destroyProgram(&prog);
        return false;
    }
    size_t cs = 0;
    n.getCUBINSize(prog, &cs);
    cubin.resize(cs);
    const bool ok = cs > 0 && n.getCUBIN(prog, cubin.data()) == NVRTC_SUCCESS;
    n.destroyProgram(&prog);
    if (!ok) err = "nvrtcGetCUBIN failed";
    if (const char* dump = getenv("MBQC_JIT_DUMP")) {  // kernel work: keep the generated source and cubin
        char name[96];
        snprintf(name, sizeof(name), "/mbqc_jit_%016llx", (unsigned long long)fnv1a(source));
        write_file_atomic(std::string(dump) + name + ".cu", std::vector<char>(source.begin(), source.end()));
        if (ok) write_file_atomic(std::string(dump) + name + ".cubin", cubin);
    }
    return ok;
}

// kernel for (plan, variant) on the current device, or nullptr (reason in state().last_error)
cudaKernel_t build_kernel(const mbqc_plan* plan, Variant v) {
    const std::string source = make_preamble(plan, v) + kind_source(v.kind);
    const uint64_t h = fnv1a(source);
    int device = 0;
    cudaGetDevice(&device);
    JitState& st = state();
    std::lock_guard<std::mutex> lock(st.mu);
    auto it = st.kernels.find({h, device});
    if (it != st.kernels.end()) return it->second.failed ? nullptr : it->second.kernel;
    Loaded ld;
    std::vector<char> cubin;
    const std::string dir = cache_dir();
    char name[64];
    snprintf(name, sizeof(name), "/%s_%016llx_nvrtc%d%d.cubin", kind_tag(v.kind), (unsigned long long)h,
             nvrtc().major, nvrtc().minor);
    bool have = !dir.empty() && read_file(dir + name, cubin);
    if (have) ++st.from_disk;
    if (!have) {
        std::string err;
        if (!compile_cubin(source, cubin, err)) {
            st.last_error = err;
            ++st.failures;
            ld.failed = true;
            st.kernels[{h, device}] = ld;
            return nullptr;
        }
        ++st.compiled;
        if (!dir.empty()) write_file_atomic(dir + name, cubin);
    }
    cudaError_t e = cudaLibraryLoadData(&ld.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&ld.kernel, ld.lib, kind_entry(v.kind));
    if (e != cudaSuccess) {
        cudaGetLastError();
        st.last_error = std::string("loading the specialised kernel: ") + cudaGetErrorString(e);
        ++st.failures;
        ld.failed = true;
        ld.kernel = nullptr;
    }
    st.kernels[{h, device}] = ld;
    return ld.kernel;
}

// per-plan memo of resolved kernels, so that a launch costs one small map lookup
struct PlanJit {
    std::mutex mu;
    std::map<Variant, cudaKernel_t> kernels;  // nullptr: tried and failed, do not retry
};
std::mutex g_plan_jit_mu;

cudaKernel_t get_kernel(const mbqc_plan* plan, Variant v) {
    PlanJit* pj = static_cast<PlanJit*>(plan->jit);
    if (!pj) {
        std::lock_guard<std::mutex> lock(g_plan_jit_mu);
        if (!plan->jit) plan->jit = new PlanJit();
        pj = static_cast<PlanJit*>(plan->jit);
    }
    std::lock_guard<std::mutex> lock(pj->mu);
    auto it = pj->kernels.find(v);
    if (it != pj->kernels.end()) return it->second;
    cudaKernel_t k = build_kernel(plan, v);
    pj->kernels[v] = k;
    return k;
}

struct JitArgsHost {  // mirrors JitArgs of sv_jit_src.inc
    const double* angles;
    double2* out;
    int* status;
    int* status_any;
    const double2* inputs;
    const double2* trig;
    long long batch;
    int input_mode;
};

const double2* trig_table_device() {
    const double2* p = nullptr;
    cudaGetSymbolAddress((void**)&p, kTrigTable128);
    return p;
}

std::atomic<int>& jit_mode_ref() {  // 0 off, 1 by batch size, 2 always
    static std::atomic<int> mode([] {
        const char* e = getenv("MBQC_JIT");
        if (!e || !*e) return 1;
        if (!strcmp(e, "0") || !strcmp(e, "off")) return 0;
        if (!strcmp(e, "force")) return 2;
        return 1;
    }());
    return mode;
}
int jit_mode() { return jit_mode_ref().load(std::memory_order_relaxed); }

}  // namespace

void mbqc_jit_free(mbqc_plan* plan) {
    delete static_cast<PlanJit*>(plan->jit);
    plan->jit = nullptr;
}

// 0 = not taken (caller continues with the ahead-of-time kernels), 1 = launched (*rc holds the result)
int mbqc_jit_try_launch(const SvBatchParams& p, const mbqc_plan* plan, int out_mode, cudaStream_t st, int64_t call_batch,
                        int* rc) {
    static const long long min_batch = [] {
        const char* e = getenv("MBQC_JIT_MIN_BATCH");
        return (e && *e) ? atoll(e) : 16384ll;
    }();
    static const int cta_env = [] {
        const char* e = getenv("MBQC_LEAN_CTA");
        return (e && *e) ? atoi(e) : 0;
    }();
    const int mode = jit_mode();
    if (mode == 0 || !plan->lean) return 0;
    if (mode == 1 && call_batch < min_batch) return 0;
    const int T = p.tab.n_angles;
    if (p.stride != T || ((uintptr_t)p.angles & 15u)) return 0;
    Variant v{out_mode, 128};
    if (out_mode == MBQC_LEAN_OUT_DIRECT && cta_env == 64) v.cta = 64;
    size_t smem = (size_t)v.cta * T * sizeof(double);
    if (out_mode != MBQC_LEAN_OUT_DIRECT) {
        const size_t stage = ((size_t)v.cta << p.tab.n_out) * sizeof(double2);
        if (stage > smem) smem = stage;
    }
    if (smem > 96 * 1024) return 0;
    cudaKernel_t kern = get_kernel(plan, v);
    if (!kern) return 0;
    if (smem > 40 * 1024) {  // static shared memory (tables, barriers) counts against the 48 KB default too
        cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            *rc = mbqc_cuda_error(e, "cudaFuncSetAttribute(mbqc_jit_sv)");
            return 1;
        }
    }
    JitArgsHost a;
    a.angles = p.angles;
    a.out = p.out;
    a.status = p.status;
    a.status_any = p.status_any;
    a.inputs = p.inputs;
    a.trig = trig_table_device();
    a.batch = p.batch;
    a.input_mode = p.input_mode;
    void* args[] = {&a};
    const unsigned blocks = (unsigned)((p.batch + v.cta - 1) / v.cta);
    cudaError_t e = cudaLaunchKernel((const void*)kern, dim3(blocks), dim3(v.cta), args, smem, st);
    if (e != cudaSuccess) {
        *rc = mbqc_cuda_error(e, "cudaLaunchKernel(mbqc_jit_sv)");
        return 1;
    }
    *rc = mbqc_after_launch("mbqc_jit_sv");
    return 1;
}

struct JitGradArgsHost {  // mirrors JitGradArgs of sv_jit_grad_src.inc
    const double* angles;
    const double2* inputs;
    const double2* target;
    const double2* trig;
    double* grad;
    double* cost;
    int* status;
    long long stride, batch, data_count;
    int input_mode;
    double gr, gi, inv2s;
    double* dst[8];
    long long row0;
    int n_dst;
};

// Parameter-shift gradient through the specialised kernel: same return convention as above.
int mbqc_jit_grad_try_launch(const mbqc_plan* plan, const SvBatchParams& p, cudaStream_t st, int* rc) {
    static const long long min_batch = [] {
        const char* e = getenv("MBQC_JIT_MIN_BATCH");
        return (e && *e) ? atoll(e) : 16384ll;
    }();
    const int mode = jit_mode();
    if (mode == 0 || !plan->lean) return 0;
    if (mode == 1 && p.batch < min_batch) return 0;
    const int T = p.tab.n_angles;
    static const int cta_env = [] {
        const char* e = getenv("MBQC_GRAD_CTA");  // 64 | 128 (kernel work)
        return (e && atoi(e) == 64) ? 64 : 128;
    }();
    static const int push_mode = [] {
        const char* e = getenv("MBQC_PUSH_MODE");  // stores | tma | tma_read (kernel work)
        if (e && !strcmp(e, "stores")) return 1;
        if (e && !strcmp(e, "tma")) return 2;
        if (e && !strcmp(e, "tma_read")) return 3;
        return 0;
    }();
    const bool push = p.push_n > 0;
    // measured (profiles/r02_c4_push_modes.jsonl, exposed time over the local-only kernel): one bulk copy
    // per GPU wins up to 4 GPUs (+13 / +24 us vs +21 / +36 us at 2 / 4 GPUs), plain stores at 8 (+72 vs +90 us)
    const int pmode = p.push_multicast ? 4 : (push_mode ? push_mode : (p.push_n > 4 ? 1 : 3));
    Variant v{push ? pmode : 0, cta_env, kKindGrad};
    const size_t smem = (size_t)v.cta * ((size_t)T + ((size_t)1 << p.tab.n_out) + ((size_t)1 << (p.tab.window - 1))) * sizeof(double2);
    if (smem > 200 * 1024 || T > 64) return 0;
    if (push) {  // the gradient is staged in the (cos, sin) slots: every column must belong to exactly one measurement
        std::vector<int> uses(T, 0);
        for (int m = 0; m < p.tab.n_steps; ++m)
            if (plan->h_steps[m].angle_idx >= 0) ++uses[plan->h_steps[m].angle_idx];
        for (int c = 0; c < T; ++c)
            if (uses[c] != 1) return 0;
        // row-major staging of the CTA's rows in the dead output / prefix regions before the bulk copies
        if (2 * (((size_t)1 << p.tab.n_out) + ((size_t)1 << (p.tab.window - 1))) < (size_t)T) return 0;
    }
    cudaKernel_t kern = get_kernel(plan, v);
    if (!kern) return 0;
    if (smem > 40 * 1024) {  // static shared memory (tables, barriers) counts against the 48 KB default too
        cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            *rc = mbqc_cuda_error(e, "cudaFuncSetAttribute(mbqc_jit_grad)");
            return 1;
        }
    }
    JitGradArgsHost a;
    a.angles = p.angles;
    a.inputs = p.inputs;
    a.target = p.target;
    a.trig = trig_table_device();
    a.grad = p.grad;
    a.cost = p.cost;
    a.status = p.status;
    a.stride = p.stride;
    a.batch = p.batch;
    a.data_count = p.data_count;
    a.input_mode = p.input_mode;
    const double sh = sin(0.5 * p.shift);
    a.gr = -2.0 * sh * sh;  // cos s - 1 without cancellation (finite differences use s = 1e-5)
    a.gi = -sin(p.shift);
    a.inv2s = 1.0 / (2.0 * p.shift);
    for (int d = 0; d < 8; ++d) a.dst[d] = d < p.push_n ? p.push_dst[d] : nullptr;
    a.row0 = p.push_row0;
    a.n_dst = p.push_n;
    void* args[] = {&a};
    const unsigned blocks = (unsigned)((p.batch + v.cta - 1) / v.cta);
    cudaError_t e = cudaLaunchKernel((const void*)kern, dim3(blocks), dim3(v.cta), args, smem, st);
    if (e != cudaSuccess) {
        *rc = mbqc_cuda_error(e, "cudaLaunchKernel(mbqc_jit_grad)");
        return 1;
    }
    *rc = mbqc_after_launch("mbqc_jit_grad");
    return 1;
}

struct DmJitArgsHost {  // mirrors DmJitArgs of dm_jit_src.inc
    const double* angles;
    long long stride;
    const double2* inputs;
    double2* out;
    signed char* outcomes;
    int* status;
    long long batch;
    int input_mode;
};

// Batched density-matrix run through the specialised kernel: same return convention as above.
int mbqc_jit_dm_try_launch(const mbqc_plan* plan, const DmBatchParams& p, cudaStream_t st, int* rc) {
    static const long long min_batch = [] {
        const char* e = getenv("MBQC_JIT_DM_MIN_BATCH");
        return (e && *e) ? atoll(e) : 1024ll;
    }();
    const int mode = jit_mode();
    if (mode == 0 || p.expect) return 0;
    if (mode == 1 && p.batch < min_batch) return 0;
    static const int lb_env = [] {
        const char* e = getenv("MBQC_DM_JIT_LB");  // register slots per lane (kernel work)
        return (e && *e) ? atoi(e) : 0;
    }();
    // window 4: 16 lanes x 4 entries per sample keep small batches short (5.8 vs 9.1 us at 1,024 samples);
    // from 4,096 samples on 4 lanes x 16 entries win (half the shared-memory traffic per step:
    // 340 vs 595 us at 262,144) -- profiles/README.md
    const int lb = lb_env ? lb_env : ((plan->tab.window == 4 && p.batch >= 4096) ? 2 : 0);
    DmJitShape sh;
    if (!dm_jit_shape(plan, sh, lb)) return 0;
    cudaKernel_t kern = get_kernel(plan, Variant{sh.lb, sh.cta, kKindDm});
    if (!kern) return 0;
    if (sh.smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
        if (e != cudaSuccess) {
            *rc = mbqc_cuda_error(e, "cudaFuncSetAttribute(mbqc_jit_dm)");
            return 1;
        }
    }
    DmJitArgsHost a;
    a.angles = p.angles;
    a.stride = p.stride;
    a.inputs = p.inputs;
    a.out = p.out;
    a.outcomes = (signed char*)p.outcomes;
    a.status = p.status;
    a.batch = p.batch;
    a.input_mode = p.input_mode;
    void* args[] = {&a};
    const int spb = sh.samples_per_cta();
    const unsigned blocks = (unsigned)((p.batch + spb - 1) / spb);
    cudaError_t e = cudaLaunchKernel((const void*)kern, dim3(blocks), dim3(sh.cta), args, sh.smem, st);
    if (e != cudaSuccess) {
        *rc = mbqc_cuda_error(e, "cudaLaunchKernel(mbqc_jit_dm)");
        return 1;
    }
    *rc = mbqc_after_launch("mbqc_jit_dm");
    return 1;
}

extern "C" {

// Human-readable state of the run-time specialisation (tests, bench): where NVRTC was found, how
// many kernels were compiled / loaded from the file cache, the last error.
const char* mbqc_jit_info(void) {
    static thread_local char buf[2048];
    Nvrtc& n = nvrtc();
    JitState& st = state();
    std::lock_guard<std::mutex> lock(st.mu);
    snprintf(buf, sizeof(buf), "mode=%d nvrtc=%s version=%d.%d compiled=%lld from_disk=%lld failures=%lld last_error=%s", jit_mode(),
             n.ok() ? n.where.c_str() : n.error.c_str(), n.major, n.minor, st.compiled, st.from_disk, st.failures,
             st.last_error.c_str());
    return buf;
}

// 0 = never specialise, 1 = calls with batch >= MBQC_JIT_MIN_BATCH, 2 = every eligible call; returns
// the previous mode (initial value from the MBQC_JIT environment variable).
int32_t mbqc_jit_set_mode(int32_t mode) {
    if (mode < 0 || mode > 2) return jit_mode();
    return jit_mode_ref().exchange(mode);
}

// Generate and compile the specialised kernel of `plan` for one output form without launching it
// (works without a GPU: NVRTC compiles offline).  Returns the cubin size, 0 when the plan is
// outside the specialised kernel's scope, or a negative MBQC_E_* code (message in mbqc_last_error).
int64_t mbqc_jit_compile_check(const mbqc_plan* plan, int32_t out_form, int32_t cta) {
    if (!plan) return mbqc_set_error(MBQC_E_ARG, "plan is NULL");
    // out_form: MBQC_OUT_SV / MBQC_OUT_DM -> mbqc_jit_sv; 100 -> the gradient kernel mbqc_jit_grad;
    // 200 -> the density-matrix kernel mbqc_jit_dm
    Variant v{out_form == MBQC_OUT_DM ? MBQC_LEAN_OUT_DM : MBQC_LEAN_OUT_DIRECT, cta == 64 ? 64 : 128};
    if (out_form == 100) v = Variant{0, 128, kKindGrad};
    if (out_form >= 101 && out_form <= 104) v = Variant{out_form - 100, 128, kKindGrad};  // replicated-result forms of the gradient kernel
    if (out_form == 200) {
        DmJitShape sh;
        if (!dm_jit_shape(plan, sh, cta == 1 || cta == 2 ? cta : 0)) return 0;  // cta 1 / 2: register slots per lane
        v = Variant{sh.lb, sh.cta, kKindDm};
    } else if (!plan->lean) {
        return 0;
    }
    const std::string source = make_preamble(plan, v) + kind_source(v.kind);
    std::vector<char> cubin;
    std::string err;
    if (!compile_cubin(source, cubin, err)) return mbqc_set_error(MBQC_E_UNSUPPORTED, err.c_str());
    return (int64_t)cubin.size();
}

}  // extern "C"
