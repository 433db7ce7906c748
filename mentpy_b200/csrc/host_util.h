// Host-side helpers shared by the translation units of the C-ABI library.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "sv_params.cuh"

// error bookkeeping of mbqc_b200.cu (thread-local message behind mbqc_last_error)
int mbqc_set_error(int code, const char* msg);
int mbqc_cuda_error(cudaError_t e, const char* what);
int mbqc_after_launch(const char* name);  // counts the launch, maps cudaGetLastError

// ---- lean register kernel (sv_lean.cuh, sv_lean_host.cu) ----
#define MBQC_LEAN_OUT_DIRECT 0  // [B][2^k] amplitudes straight from registers
#define MBQC_LEAN_OUT_STAGED 1  // the same, written CTA-coalesced (page-locked host output)
#define MBQC_LEAN_OUT_DM 2      // [B][4^k] |psi><psi|
void mbqc_lean_build_proto(mbqc_plan* plan);  // sets plan->lean (or leaves it null)
void mbqc_lean_free_proto(mbqc_plan* plan);
// 0 = not eligible (caller falls back to sv_reg_kernel), 1 = launched (*rc holds the result)
int mbqc_lean_try_launch(const mbqc::SvBatchParams& p, const mbqc_plan* plan, int out_mode, cudaStream_t st, int* rc);

// ---- run-time specialised kernel (sv_jit_src.inc, sv_jit_host.cu): same return convention ----
void mbqc_jit_free(mbqc_plan* plan);
int mbqc_jit_try_launch(const mbqc::SvBatchParams& p, const mbqc_plan* plan, int out_mode, cudaStream_t st,
                        int64_t call_batch, int* rc);
int mbqc_jit_grad_try_launch(const mbqc_plan* plan, const mbqc::SvBatchParams& p, cudaStream_t st, int* rc);
namespace mbqc {
struct DmBatchParams;  // dm_batch.cuh
}
int mbqc_jit_dm_try_launch(const mbqc_plan* plan, const mbqc::DmBatchParams& p, cudaStream_t st, int* rc);
