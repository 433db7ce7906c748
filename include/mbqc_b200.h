/*
 * mbqc_b200.h -- C ABI of the B200-native MBQC pattern-simulation hot path.
 *
 * The reference (bestquark/mentpy) is pure Python and has no FFI; its boundary for this path is the
 * `BaseSimulator` ABC (mentpy/simulators/base_simulator.py:13-101) behind the `PatternSimulator`
 * facade (mentpy/simulators/pattern_simulator.py:38-88).  Each entry point below names the
 * reference method(s) whose arithmetic it replaces.  The Python backend classes
 * (mentpy_b200/simulators/) bind these with ctypes; INTEGRATION.md shows the stub a mentpy
 * maintainer would add.
 *
 * Conventions
 *   - every call returns 0 on success, a negative MBQC_E_* code on failure; mbqc_last_error()
 *     returns a thread-local message for the last failure.
 *   - the caller owns every device buffer; the library owns only the opaque, immutable plan
 *     (a few hundred bytes of device memory).  No hidden allocation in any run/step call.
 *   - every run/step call is asynchronous on the `stream` argument (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream).
 *   - complex numbers are interleaved (re, im) pairs of the plan's real type (double by default).
 *   - a "slot" is a bit position of the state index (0 = least significant).  Qubits are never
 *     shifted: the qubit appended after a measurement re-uses the measured qubit's slot.
 */
#ifndef MBQC_B200_H
#define MBQC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBQC_OK 0
#define MBQC_E_ARG (-1)      /* invalid argument (maps to ValueError) */
#define MBQC_E_CUDA (-2)     /* CUDA runtime / launch failure (maps to RuntimeError) */
#define MBQC_E_UNSUPPORTED (-3) /* window / size outside what the kernels cover */

/* measurement planes after host lowering: X and Y are fixed-angle XY (ment.py:233-235) */
#define MBQC_PLANE_XY 0
#define MBQC_PLANE_XZ 1
#define MBQC_PLANE_YZ 2
#define MBQC_PLANE_Z 3 /* DM path, expectation mode only: trace the qubit out, record prob1 */
#define MBQC_PLANE_XYZ 4 /* DM path, fixed angles only (ment.py:239-251): M = fixed_cos X + fixed_sin Y + fixed_z Z */

#define MBQC_STEP_APPEND 1u /* a |+> qubit enters the freed slot and is CZ'ed with nbr_mask */

/* per-sample status bits written by the batched kernels */
#define MBQC_STATUS_OK 0
#define MBQC_STATUS_BAD_NORM 1   /* zero / non-finite norm or trace (np_simulator_dm.py:179-182) */
#define MBQC_STATUS_OUTCOME1 2   /* DM: some step took outcome 1 (np_simulator_dm.py:335-338) */

#define MBQC_OUT_SV 0 /* [B][2^k] amplitudes            (np_simulator_sv.py:294-295) */
#define MBQC_OUT_DM 1 /* [B][2^k][2^k] density matrices (np_simulator_sv.py:292-293) */

#define MBQC_INPUT_PLUS 0   /* |+>^{|I|}                    (pattern_simulator.py:58-61) */
#define MBQC_INPUT_SHARED 1 /* one [2^|I|] state for all samples */
#define MBQC_INPUT_BATCH 2  /* [B][2^|I|], one state per sample */

#define MBQC_MAX_WINDOW_REG 5    /* state-vector window held in registers, one thread per sample */
#define MBQC_MAX_WINDOW_SMEM_SV 12
#define MBQC_MAX_WINDOW_SMEM_DM 6
#define MBQC_MAX_WINDOW 40       /* streaming / sharded regime (slots fit a uint64 mask) */

/* One measurement of the pattern = one iteration of NumpySimulatorSV.measure
 * (np_simulator_sv.py:164-225) / NumpySimulatorDM.measure (np_simulator_dm.py:151-216). */
typedef struct mbqc_step {
    int32_t slot;       /* slot of the measured qubit */
    int32_t angle_idx;  /* >= 0: column of the angle matrix; -1: fixed angle below */
    int32_t plane;      /* MBQC_PLANE_* (SV path accepts XY only, np_simulator_sv.py:54-59) */
    uint32_t flags;     /* MBQC_STEP_APPEND */
    double fixed_cos;   /* cos/sin of the fixed angle, evaluated on the host exactly as the */
    double fixed_sin;   /*   reference does (np.cos/np.sin, or exact 1,0 / 0,1 for planes X / Y) */
    uint64_t nbr_mask;  /* slots of the appended qubit's in-window neighbours */
    double fixed_z;     /* MBQC_PLANE_XYZ: Z component sin(t2); (fixed_cos, fixed_sin) = cos(t2) (cos t1, sin t1) */
    /* Outcome-controlled measurement (operators/controlled_ment.py:14-113), density-matrix path:
     * cond_mask != 0 selects earlier outcomes (bit j = the step j+1 measurements back, at most 32);
     * their values, packed lowest selected bit first, index cond_table; a 1 there replaces
     * (plane, angle_idx, fixed_*) by the alt_* fields for that sample.  cond_mask == 0: plain step. */
    uint32_t cond_mask;
    uint32_t cond_table;
    int32_t alt_plane;
    int32_t alt_angle_idx;
    double alt_cos, alt_sin, alt_z;
} mbqc_step;

/* Single-qubit channel in block form on (rho00, rho01, rho10, rho11) of the affected qubit:
 *   rho00' = pop[0] rho00 + pop[1] rho11      rho01' = coh_g rho01 + coh_d rho10
 *   rho11' = pop[2] rho00 + pop[3] rho11      rho10' = coh_g rho10 + coh_d rho01
 * Covers depolarizing, phase/bit flip, amplitude / phase damping and generalized amplitude
 * damping (the channel list of mentpy/simulators/pennylane_simulator.py:123-136).  Applied to
 * every measured qubit right before its measurement and to every output qubit at the end. */
typedef struct mbqc_noise {
    double pop[4];
    double coh_g;
    double coh_d;
} mbqc_noise;

/* Flow corrections of one measurement step for sampled runs (force0 = False; the reference raises
 * NotImplementedError, np_simulator_sv.py:50-51; rule: pennylane_simulator.py:145-153).  The
 * kernels keep the last 32 outcomes of a sample in a shift register (bit d = outcome of step
 * m-1-d): xdep / zdep select the outcomes that put a pending X / Z on the qubit measured at this
 * step (its XY angle becomes (-1)^a theta + b pi); outx / outz name the output qubits (bit q =
 * q-th output) whose X / Z byproduct toggles when THIS step's outcome is 1. */
typedef struct mbqc_feedforward {
    uint32_t xdep, zdep, outx, outz;
} mbqc_feedforward;

#define MBQC_OUTCOMES_SAMPLE 0 /* draw outcomes from the Born rule (Philox4x32-10, stateless) */
#define MBQC_OUTCOMES_FORCED 1 /* read the outcome record from d_outcomes */

typedef struct mbqc_plan mbqc_plan;

/* Lowered pattern.  Replaces the per-run bookkeeping of NumpySimulatorSV.__init__/reset
 * (np_simulator_sv.py:38-128, :299-320: window seeding, cached initial CZ product) and the
 * output reordering (np_simulator_sv.py:286-290, np_simulator_dm.py:267-273).
 *   input_slot[q]   slot of input qubit q (q = 0 is the MSB of the caller's input index)
 *   init_cz_mask[a] for slot a: mask of slots b > a CZ'ed with a in the first window
 *   output_slot[q]  slot of output qubit q (q = 0 is the MSB of the output index)
 *   noise           NULL for the noiseless path */
int mbqc_plan_create(const mbqc_step* steps, int32_t n_steps, int32_t window, int32_t n_inputs,
                     int32_t n_outputs, int32_t n_angles, const int32_t* input_slot,
                     const uint64_t* init_cz_mask, const int32_t* output_slot,
                     const mbqc_noise* noise, mbqc_plan** out);
void mbqc_plan_destroy(mbqc_plan* plan);

/* Same lowering tables WITHOUT any device allocation (no CUDA call): such a plan cannot be run; it
 * exists so that the run-time specialised kernel of a pattern can be generated and compiled where
 * there is no GPU (mbqc_jit_compile_check: build boxes, the CPU test suite). */
int mbqc_plan_create_hostonly(const mbqc_step* steps, int32_t n_steps, int32_t window, int32_t n_inputs,
                              int32_t n_outputs, int32_t n_angles, const int32_t* input_slot,
                              const uint64_t* init_cz_mask, const int32_t* output_slot,
                              const mbqc_noise* noise, mbqc_plan** out);

/* Run-time specialisation of mbqc_run_batch_sv (window <= MBQC_MAX_WINDOW_REG, reference schedules):
 * for calls with batch >= 16384 (MBQC_JIT_MIN_BATCH; MBQC_JIT=0 disables, MBQC_JIT=force always)
 * the library generates the pattern's kernel with its angle columns, CZ signs and step count as
 * compile-time constants, compiles it once with NVRTC (dlopen'ed; cubins are cached under
 * MBQC_JIT_CACHE, default ~/.cache/mentpy_b200) and launches that instead of the general kernels.
 * Without NVRTC the general CUDA kernels run.  compile_check: generate + compile only; returns the
 * cubin size in bytes, 0 if the plan is outside the specialised kernel's scope, < 0 on error.
 * info: where NVRTC was found, kernels compiled / loaded from the file cache, last error. */
int64_t mbqc_jit_compile_check(const mbqc_plan* plan, int32_t out_form, int32_t cta);
const char* mbqc_jit_info(void);
/* 0 = never, 1 = by batch size (default), 2 = every eligible call; returns the previous mode */
int32_t mbqc_jit_set_mode(int32_t mode);

/* NumpySimulatorSV.run over a batch (np_simulator_sv.py:227-297): angles [B][T] (row stride
 * `angle_stride` doubles) -> out.  Amplitudes carry the reference's global phase
 * prod_j (1+e^{i th_j})/|1+e^{i th_j}| so that 'sv' outputs compare amplitude by amplitude.
 */
int mbqc_run_batch_sv(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                      const void* d_inputs, int32_t input_mode, int64_t batch, void* d_out,
                      int32_t out_form, int32_t* d_status, void* stream);

/* complex64 mode of mbqc_run_batch_sv (window <= MBQC_MAX_WINDOW_REG): fp32 state and arithmetic,
 * d_inputs / d_out hold interleaved float pairs; angles stay fp64.  Accuracy target: infidelity
 * <= 1e-5 against the fp64 path (BASELINE.json north_star). */
int mbqc_run_batch_sv_f32(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                          const void* d_inputs, int32_t input_mode, int64_t batch, void* d_out,
                          int32_t out_form, int32_t* d_status, void* stream);

/* Host-buffer form of mbqc_run_batch_sv -- the end-to-end plugin call: h_angles [B][T] and h_out
 * live in host memory (page-locked for full PCIe speed).  The batch is cut into n_chunks pieces
 * (<= 0: library default) whose H2D copy, kernel and D2H copy are queued on internal streams so
 * that transfers in both directions overlap each other and the kernels.  d_work is caller-owned
 * device scratch of at least mbqc_host_workspace_bytes() bytes.  Synchronous: h_out is complete
 * on return.  h_status_any (may be NULL) receives the OR of all per-sample status bits. */
int64_t mbqc_host_workspace_bytes(const mbqc_plan* plan, int64_t batch, int32_t out_form);
int mbqc_run_batch_sv_host(const mbqc_plan* plan, const double* h_angles, int64_t angle_stride,
                           const void* d_inputs, int32_t input_mode, int64_t batch, void* h_out,
                           int32_t out_form, void* d_work, int64_t work_bytes,
                           int32_t* h_status_any, int32_t n_chunks);

/* Asynchronous form: _submit queues the call and returns a ticket at once, mbqc_host_wait blocks
 * until that call's h_out is complete.  Up to 4 calls may be in flight per device (each with its
 * own h_angles / h_out / d_work); consecutive calls run on alternating stream sets, so the H2D
 * copies of call n+1 overlap the kernels and result transfers of call n (PCIe is full duplex).
 * mbqc_run_batch_sv_host == submit + wait. */
int mbqc_run_batch_sv_host_submit(const mbqc_plan* plan, const double* h_angles, int64_t angle_stride,
                                  const void* d_inputs, int32_t input_mode, int64_t batch, void* h_out,
                                  int32_t out_form, void* d_work, int64_t work_bytes, int32_t n_chunks,
                                  int32_t* ticket);
int mbqc_host_wait(int32_t ticket, int32_t* h_status_any);

/* NumpySimulatorDM.run over a batch (np_simulator_dm.py:218-283), optional noise from the plan:
 * out [B][2^k][2^k]; d_outcomes (may be NULL) [B][n_steps] int8 receives the outcome record
 * (simulator.outcomes). */
int mbqc_run_batch_dm(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                      const void* d_inputs, int32_t input_mode, int64_t batch, void* d_out,
                      int8_t* d_outcomes, int32_t* d_status, void* stream);

/* mode="expectation" of NumpySimulatorDM.run (np_simulator_dm.py:327-344): as mbqc_run_batch_dm,
 * but steps in MBQC_PLANE_Z are not projected -- the qubit is traced out and prob1 / (prob0 + prob1)
 * is recorded in d_expect[b][step] (double, [B][n_steps], 0 for the other steps; required when
 * the plan has plane-Z steps) -- the classifier read-out of a pattern.  Window <= 5.  Plans with
 * plane-Z steps are rejected by mbqc_run_batch_dm (the reference draws those outcomes at random
 * there, even under force0). */
int mbqc_run_batch_dm_expect(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                             const void* d_inputs, int32_t input_mode, int64_t batch, void* d_out,
                             int8_t* d_outcomes, double* d_expect, int32_t* d_status, void* stream);

/* Attach the feed-forward table (n_steps records) to a plan.  Call once, right after
 * mbqc_plan_create and before the plan is shared between threads. */
int mbqc_plan_set_feedforward(mbqc_plan* plan, const mbqc_feedforward* ff, int32_t n_steps);

/* Sampled twins of mbqc_run_batch_sv / _dm (window <= 5, XY-plane steps): every measurement takes
 * outcome 0 with its Born probability -- uniform draw = Philox4x32-10 with key `seed` and counter
 * (sample_offset + b, step), so a sample's record does not depend on how the batch is split -- or
 * the outcome given in d_outcomes (MBQC_OUTCOMES_FORCED); later angles adapt through the plan's
 * feed-forward table; with correct != 0 the byproduct X^a Z^b is applied to the output register,
 * after which a noiseless sample equals the deterministic (force0) state up to a global phase.
 * d_out: [B][2^k] (sv) / [B][2^k][2^k] (dm) complex128.  Optional outputs (NULL to skip):
 * d_outcomes [B][n_steps] int8 (input when forced), d_byproducts [B] (x bits | z bits << 16),
 * d_prob [B] probability of the record, d_status [B]. */
int mbqc_run_batch_sv_sampled(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                              const void* d_inputs, int32_t input_mode, int64_t batch, uint64_t seed,
                              uint64_t sample_offset, int32_t outcome_mode, int32_t correct, void* d_out,
                              int8_t* d_outcomes, uint32_t* d_byproducts, double* d_prob, int32_t* d_status,
                              void* stream);
int mbqc_run_batch_dm_sampled(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                              const void* d_inputs, int32_t input_mode, int64_t batch, uint64_t seed,
                              uint64_t sample_offset, int32_t outcome_mode, int32_t correct, void* d_out,
                              int8_t* d_outcomes, uint32_t* d_byproducts, double* d_prob, int32_t* d_status,
                              void* stream);

/* mbqc_run_batch_dm for patterns with plane-Z steps in the reference's mode="sample"
 * (np_simulator_dm.py:329-346): even under force0 the outcome of a plane-Z step is drawn from
 * (prob0, prob1) -- here from the Philox stream (seed, sample_offset + b, step), so a run is
 * reproducible -- and the state is projected on |0><0| or |1><1|; every other step follows the
 * deterministic rule.  d_outcomes [B][n_steps] (may be NULL) receives the record. */
int mbqc_run_batch_dm_zsample(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                              const void* d_inputs, int32_t input_mode, int64_t batch, uint64_t seed,
                              uint64_t sample_offset, void* d_out, int8_t* d_outcomes, int32_t* d_status, void* stream);

/* Batched parameter-shift / central-difference gradient of cost(x) = 1 - |<target|psi(x)>|^2:
 * grad[b][i] = (cost(x_b + s e_i) - cost(x_b - s e_i)) / (2 s)
 * (gradients/_parameter_shift.py:9-25 with s = 1.5; _finite_difference.py:9-25 central with
 * s = h).  d_cost (may be NULL) [B] receives cost(x_b).  d_target [2^k] complex. */
int mbqc_psr_grad_batch(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                        const void* d_inputs, int32_t input_mode, int64_t batch,
                        const void* d_target, double shift, double* d_grad, double* d_cost,
                        int32_t* d_status, void* stream);

/* The same gradient with a REPLICATED result: the `batch` rows of this call are rows
 * first_row .. first_row + batch of a [total][T] float64 result that lives on n_results GPUs;
 * d_results[0] is the local copy, d_results[1..] are the other GPUs' copies (peer-mapped: CUDA IPC,
 * mbqc_ipc_import).  The kernel stores each finished tile of rows into every copy (NVLink peer
 * stores overlap the remaining computation): the multi-GPU form of gradients/_parameter_shift.py:9-25
 * where the final gather is part of the launch.  Follow with mbqc_peer_barrier before reading. */
int mbqc_psr_grad_batch_push(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                             const void* d_inputs, int32_t input_mode, int64_t batch,
                             const void* d_target, double shift, void* const* d_results, int32_t n_results,
                             int64_t first_row, double* d_cost, int32_t* d_status, void* stream);

/* The replicated result through ONE NVSwitch multicast address (cuMulticast* / torch symmetric memory):
 * every store is replicated by the switch into all GPUs' copies, so a GPU sends its rows once
 * instead of once per peer.  MBQC_E_UNSUPPORTED when the specialised kernel is not available
 * (multicast addresses accept multimem stores only): use mbqc_psr_grad_batch_push then. */
int mbqc_psr_grad_batch_multicast(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                                  const void* d_inputs, int32_t input_mode, int64_t batch,
                                  const void* d_target, double shift, void* d_result_multicast,
                                  int64_t first_row, double* d_cost, int32_t* d_status, void* stream);

/* Data-set averaged cost and its gradient in one call -- the S x 2T pattern evaluations of one
 * optimiser step of the reference's training loop (docs/tutorials/intro-to-mbqml.rst:35-86:
 * cost(x) = mean_s [1 - |<t_s|psi(x; in_s)>|^2], differentiated by gradients/_parameter_shift.py:9-25):
 *     d_cost[p]    = mean_s cost_s(x_p)
 *     d_grad[p][i] = mean_s (cost_s(x_p + s e_i) - cost_s(x_p - s e_i)) / (2 s)
 * d_angles [P][stride]; d_inputs [S][2^|I|] complex128 or NULL (|+> inputs); d_targets [S][2^k];
 * d_grad [P][T]; d_cost [P] / d_status [P] may be NULL.  d_workspace: caller-owned device buffer of
 * mbqc_psr_grad_dataset_workspace_bytes(plan, P, S) bytes (per-sample gradients before the mean;
 * the reduction order is fixed, so results are reproducible run to run). */
int64_t mbqc_psr_grad_dataset_workspace_bytes(const mbqc_plan* plan, int64_t n_vectors, int64_t n_data);
int mbqc_psr_grad_dataset(const mbqc_plan* plan, const double* d_angles, int64_t angle_stride,
                          const void* d_inputs, const void* d_targets, int64_t n_vectors, int64_t n_data,
                          double shift, double* d_grad, double* d_cost, int32_t* d_status,
                          void* d_workspace, void* stream);

/* Device-resident training loop: num_iters iterations of [data-set gradient -> optimiser update]
 * queued on `stream` from one call, two kernel launches per iteration, nothing returns to the
 * host in between (the reference does 2T x S simulator runs and one numpy update per iteration:
 * optimizers/adam.py:54-108, sgd.py:49-96 driven by gradients/_parameter_shift.py:9-25).
 * d_x [P][T] (contiguous) is updated in place.  d_state: caller-owned, [2][P][T] doubles (Adam m, v;
 * SGD uses the first half as velocity), zeroed by the caller before the first call;
 * first_iteration = number of updates already applied (Adam's bias correction continues from
 * there).  d_cost_history (may be NULL) [num_iters][P]: data-set cost BEFORE each update.
 * d_status (may be NULL) [P]: OR over all iterations.  d_workspace: as mbqc_psr_grad_dataset. */
#define MBQC_OPT_ADAM 1
#define MBQC_OPT_SGD 2
typedef struct mbqc_optimizer {
    int32_t kind;     /* MBQC_OPT_* */
    int32_t nesterov; /* SGD only */
    double step_size;
    double b1, b2, eps; /* Adam */
    double momentum;    /* SGD */
} mbqc_optimizer;
int mbqc_train_dataset(const mbqc_plan* plan, double* d_x, const void* d_inputs, const void* d_targets,
                       int64_t n_vectors, int64_t n_data, double shift, const mbqc_optimizer* opt,
                       int32_t first_iteration, int32_t num_iters, double* d_state, double* d_cost_history,
                       int32_t* d_status, void* d_workspace, void* stream);

/* ---- streaming regime: one large window resident in HBM, one angle set -------------------------
 * Replaces NumpySimulatorSV.reset / measure / run for windows the reference cannot hold (it
 * materialises 2^w x 2^w operators: np_simulator_sv.py:103-128, :164-225, :286-297).  The state is
 * 2^local_bits complex numbers per GPU; with G = 2^g GPUs the top g index bits are the rank
 * ("shard slots").  The host mirror (mentpy_b200/streaming.py) turns the plan's steps into pass
 * descriptors; slots in these calls are PHYSICAL bit positions of the global index. */
#define MBQC_STREAM_MAX_FUSE 5
#define MBQC_STREAM_MAX_RANGES 16

typedef struct mbqc_stream_desc {
    int32_t n_fused;         /* K consecutive measurements applied in this pass (1..5) */
    int32_t n_ranges;        /* zero fields squeezed out of the thread index (fused + dead slots) */
    uint32_t range_pos[MBQC_STREAM_MAX_RANGES];   /* ascending, in index coordinates after the */
    uint32_t range_width[MBQC_STREAM_MAX_RANGES]; /*   previous inserts                         */
    uint64_t elem_offset[MBQC_STREAM_MAX_FUSE];   /* 1 << slot of fused measurement j */
    double cos_t[MBQC_STREAM_MAX_FUSE];
    double sin_t[MBQC_STREAM_MAX_FUSE];
    uint64_t nbr_mask[MBQC_STREAM_MAX_FUSE];      /* slots of the appended qubit's neighbours */
    uint32_t local_mask[MBQC_STREAM_MAX_FUSE];    /* ... restricted to the fused slots (bit i = measurement i) */
    uint32_t append_mask;    /* bit j: measurement j appends a |+> qubit (else its slot dies) */
    uint64_t n_groups;       /* 2^(live local bits - K) */
    uint64_t index_or;       /* rank << local_bits: completes the index for sign parities */
    double scale;            /* exact power-of-two rescale applied on load */
    uint64_t elem_bit[MBQC_STREAM_MAX_FUSE];      /* 1 << slot as an INDEX bit (elem_offset is an address
                                                     offset and differs for the split top local slot) */
} mbqc_stream_desc;

/* Window seed (np_simulator_sv.py:103-128) for passes that generate the initial state on the fly. */
typedef struct mbqc_stream_seed {
    int32_t window;
    int32_t n_inputs;
    int32_t input_slot[16];
    uint64_t init_cz_mask[MBQC_MAX_WINDOW];
    const void* d_input;     /* NULL: |+> inputs */
    double scale;            /* 2^{-(w-|I|)/2} (or 2^{-w/2} with |+> inputs) */
} mbqc_stream_seed;

/* Seed the local share of the window: input (x) |+>^(w-|I|), initial CZ signs (np_simulator_sv.py:103-128). */
int mbqc_stream_init(void* d_state, int32_t local_bits, uint64_t index_or, int32_t window,
                     int32_t n_inputs, const int32_t* input_slot, const uint64_t* init_cz_mask,
                     const void* d_input, double scale, void* stream);
/* One in-place pass over the local share applying desc->n_fused measurements (np_simulator_sv.py:164-225). */
int mbqc_stream_steps(void* d_state, const mbqc_stream_desc* desc, void* stream);
/* Variant for fused slots that are all among the 5 lowest index bits: one element per thread, pair
 * partners exchanged with warp shuffles, every access a contiguous 512-byte span per warp.
 * Descriptor convention differs: ranges squeeze out DEAD slots only (all >= 5) and n_groups is the
 * number of live local elements (multiple of 32); elem_bit[j] = 1 << slot_j < 32. */
int mbqc_stream_steps_lanes(void* d_state, const mbqc_stream_desc* desc, void* stream);
/* Same pass, but the amplitudes it consumes are generated from the seed instead of being read:
 * the first pass of a pattern then needs no separate init pass (saves one write + one read of
 * the whole state). */
int mbqc_stream_steps_seeded(void* d_state, const mbqc_stream_desc* desc, const mbqc_stream_seed* seed,
                             void* stream);
/* Measurement of a shard slot fused with the NVLink transfer: d_peer is the partner GPU's half
 * (peer-mapped memory), read directly by the kernel.  role 0/1: this rank's shard bit; role 2:
 * tail step without append (survivor side).  n = live elements of the half; dead slots are squeezed
 * out of the loop index with the zero-field ranges (as in mbqc_stream_desc).  See csrc/stream.cuh. */
int mbqc_stream_exchange(void* d_own, const void* d_peer, void* d_spare, int32_t role, double cos_t,
                         double sin_t, double scale, uint64_t nbr_mask, int32_t const_parity,
                         uint64_t n, int32_t n_ranges, const uint32_t* range_pos,
                         const uint32_t* range_width, void* stream);
/* Collect the output amplitudes this rank owns into d_out [2^k] (others untouched; np_simulator_sv.py:286-297). */
int mbqc_stream_gather(const void* d_state, int32_t local_bits, uint64_t index_or,
                       int32_t n_outputs, const int32_t* output_slot, void* d_out, void* stream);

/* Plain cudaMalloc'ed buffers (IPC-exportable, unlike sub-allocations of a caching allocator) and
 * CUDA IPC handles (64 bytes) so that one process per GPU can map its neighbours' shards. */
int mbqc_device_alloc(int64_t bytes, void** d_ptr);
int mbqc_device_free(void* d_ptr);
int mbqc_ipc_export(const void* d_ptr, void* handle64);
int mbqc_ipc_import(const void* handle64, void** d_ptr);
int mbqc_ipc_close(void* d_ptr);

/* Stream-ordered barrier across the GPUs of one node through flag words in peer-mapped memory:
 * d_flags[r] = rank r's array of 16 uint64 (zeroed once; entry `rank` is the caller's own
 * allocation, the others are IPC imports; the kernel keeps its call count in word 8, so the launch
 * has no changing argument and can be captured in a CUDA graph).  Every rank makes the same
 * sequence of calls.  Work queued on `stream` after the call starts only when every rank's earlier
 * work on its stream -- including its peer stores -- has completed. */
int mbqc_peer_barrier(void* const* d_flags, int32_t n_ranks, int32_t rank, void* stream);

/* ---- calculator helpers (mentpy/calculator/state_ops.py), single state, qubit 0 = MSB ------------
 * pure: SUM over the traced qubits then renormalise (:42-74 -- the reference's pure-state "partial
 * trace"); mixed: true partial trace (:77-119); pure2density: |psi><psi| (:16-39). n_qubits <= 14. */
int mbqc_partial_trace_pure(const void* d_psi, int32_t n_qubits, const int32_t* traced, int32_t n_traced,
                            void* d_out, void* stream);
int mbqc_partial_trace_mixed(const void* d_rho, int32_t n_qubits, const int32_t* traced, int32_t n_traced,
                             void* d_out, void* stream);
int mbqc_pure2density(const void* d_psi, int32_t n_qubits, void* d_out, void* stream);

/* plan introspection (used by the host mirror and the tests) */
int32_t mbqc_plan_window(const mbqc_plan* plan);
int32_t mbqc_plan_num_steps(const mbqc_plan* plan);
int32_t mbqc_plan_num_outputs(const mbqc_plan* plan);

/* ---- measurement probes (bench.py's second roofline; not on the product path) -------------------
 * fp64_fma: `blocks` x `threads` (<= 256) threads run `iters` rounds of 8 independent DFMAs each;
 * *flops receives the flop count of the launch (2 per FMA).  copy: grid-stride 16-byte copy of
 * `bytes` (multiple of 16) from d_src to d_dst -- HBM traffic 2 * bytes. */
int mbqc_probe_fp64_fma(int64_t iters, int32_t blocks, int32_t threads, double* d_out, int64_t* flops, void* stream);
int mbqc_probe_copy(const void* d_src, void* d_dst, int64_t bytes, int32_t blocks, void* stream);

/* number of kernels this library has launched in the calling process (bench gpu_launches) */
int64_t mbqc_launch_count(void);

const char* mbqc_last_error(void);
const char* mbqc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MBQC_B200_H */
