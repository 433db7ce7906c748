// Parameter block of the batched density-matrix kernels (dm_batch.cuh, dm_reg.cuh, dm_jit_src.inc).
#pragma once
#include "common.cuh"

namespace mbqc {

struct DmBatchParams {
    PlanTables tab;
    const StepDev* __restrict__ steps;
    const double* __restrict__ angles;
    int64_t stride;
    const double2* __restrict__ inputs;
    int32_t input_mode;
    int64_t batch;
    double2* __restrict__ out;      // [B][4^k]
    int8_t* __restrict__ outcomes;  // [B][n_steps] or null
    int32_t* __restrict__ status;
    double* __restrict__ expect;    // [B][n_steps] prob1 of plane-Z steps (expectation mode) or null
    // plane-Z steps in mode="sample" (np_simulator_dm.py:329-333): outcome drawn from (prob0, prob1)
    // with the Philox stream (z_seed, z_offset + sample), then projected on |0><0| / |1><1|
    int32_t z_sample;
    uint64_t z_seed, z_offset;
};

}  // namespace mbqc
