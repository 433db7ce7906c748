// Parameter block of the batched state-vector kernels (shared by every translation unit).
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
    // data-set mode of the gradient kernels (data_count = S > 0): sample b = p * S + s uses angle
    // row p, input state s and target state s; 0 = every sample has its own row / input, one target
    int64_t data_count;
    // replicated result (mbqc_psr_grad_batch_push): the rows of this launch are rows push_row0..
    // of a [total][T] result held by push_n GPUs; push_dst[0] is the local copy, the others are
    // peer-mapped.  push_n = 0: plain `grad` output.
    double* push_dst[8];
    int64_t push_row0;
    int32_t push_n;
    int32_t push_multicast;  // push_dst[0] is an NVSwitch multicast address covering every copy
};

// (angle row, data item) of sample b
__device__ __forceinline__ void sample_index(const SvBatchParams& p, int64_t b, int64_t& row, int64_t& item) {
    row = b;
    item = b;
    if (p.data_count > 0) {
        row = b / p.data_count;
        item = b - row * p.data_count;
    }
}

}  // namespace mbqc
