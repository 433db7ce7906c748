"""ctypes binding of the C ABI declared in include/mbqc_b200.h.

The shared library is built in-tree by `build.sh` / `__graft_entry__.build()` (nvcc, sm_100a).
There is deliberately NO fallback: if the library is missing or a call fails, the product raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MBQC_LIB_PATH", os.path.join(_HERE, "_mbqc_b200.so"))  # override: kernel experiments

MBQC_OK, MBQC_E_ARG, MBQC_E_CUDA, MBQC_E_UNSUPPORTED = 0, -1, -2, -3
PLANE_XY, PLANE_XZ, PLANE_YZ, PLANE_Z, PLANE_XYZ = 0, 1, 2, 3, 4
STEP_APPEND = 1
STATUS_BAD_NORM, STATUS_OUTCOME1 = 1, 2
OUT_SV, OUT_DM = 0, 1
OUTCOMES_SAMPLE, OUTCOMES_FORCED = 0, 1
INPUT_PLUS, INPUT_SHARED, INPUT_BATCH = 0, 1, 2
MAX_WINDOW_REG, MAX_WINDOW_SMEM_SV, MAX_WINDOW_SMEM_DM, MAX_WINDOW = 5, 12, 6, 40
MAX_IO = 16


class Step(C.Structure):
    _fields_ = [("slot", C.c_int32), ("angle_idx", C.c_int32), ("plane", C.c_int32),
                ("flags", C.c_uint32), ("fixed_cos", C.c_double), ("fixed_sin", C.c_double),
                ("nbr_mask", C.c_uint64), ("fixed_z", C.c_double), ("cond_mask", C.c_uint32), ("cond_table", C.c_uint32),
                ("alt_plane", C.c_int32), ("alt_angle_idx", C.c_int32), ("alt_cos", C.c_double), ("alt_sin", C.c_double),
                ("alt_z", C.c_double)]


class Optimizer(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nesterov", C.c_int32), ("step_size", C.c_double), ("b1", C.c_double),
                ("b2", C.c_double), ("eps", C.c_double), ("momentum", C.c_double)]


OPT_ADAM, OPT_SGD = 1, 2


class FeedForwardC(C.Structure):
    _fields_ = [("xdep", C.c_uint32), ("zdep", C.c_uint32), ("outx", C.c_uint32), ("outz", C.c_uint32)]


STREAM_MAX_FUSE, STREAM_MAX_RANGES = 5, 16


class StreamDesc(C.Structure):
    _fields_ = [("n_fused", C.c_int32), ("n_ranges", C.c_int32),
                ("range_pos", C.c_uint32 * STREAM_MAX_RANGES), ("range_width", C.c_uint32 * STREAM_MAX_RANGES),
                ("elem_offset", C.c_uint64 * STREAM_MAX_FUSE), ("cos_t", C.c_double * STREAM_MAX_FUSE),
                ("sin_t", C.c_double * STREAM_MAX_FUSE), ("nbr_mask", C.c_uint64 * STREAM_MAX_FUSE),
                ("local_mask", C.c_uint32 * STREAM_MAX_FUSE), ("append_mask", C.c_uint32),
                ("n_groups", C.c_uint64), ("index_or", C.c_uint64), ("scale", C.c_double),
                ("elem_bit", C.c_uint64 * STREAM_MAX_FUSE)]


class StreamSeed(C.Structure):
    _fields_ = [("window", C.c_int32), ("n_inputs", C.c_int32), ("input_slot", C.c_int32 * 16),
                ("init_cz_mask", C.c_uint64 * MAX_WINDOW), ("d_input", C.c_void_p), ("scale", C.c_double)]


class Noise(C.Structure):
    _fields_ = [("pop", C.c_double * 4), ("coh_g", C.c_double), ("coh_d", C.c_double)]


_SIGNATURES = {
    "mbqc_plan_create": (C.c_int, [C.POINTER(Step), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_int32), C.POINTER(Noise), C.POINTER(C.c_void_p)]),
    "mbqc_plan_create_hostonly": (C.c_int, [C.POINTER(Step), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_int32), C.POINTER(Noise), C.POINTER(C.c_void_p)]),
    "mbqc_jit_compile_check": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32]),
    "mbqc_jit_info": (C.c_char_p, []),
    "mbqc_jit_set_mode": (C.c_int32, [C.c_int32]),
    "mbqc_plan_destroy": (None, [C.c_void_p]),
    "mbqc_run_batch_sv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                    C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "mbqc_run_batch_sv_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                        C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "mbqc_host_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int32]),
    "mbqc_run_batch_sv_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                         C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                         C.POINTER(C.c_int32), C.c_int32]),
    "mbqc_run_batch_sv_host_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                                C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                                C.c_int32, C.POINTER(C.c_int32)]),
    "mbqc_host_wait": (C.c_int, [C.c_int32, C.POINTER(C.c_int32)]),
    "mbqc_run_batch_dm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                    C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mbqc_psr_grad_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                      C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "mbqc_run_batch_dm_zsample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
                                            C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mbqc_psr_grad_batch_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                           C.c_int64, C.c_void_p, C.c_double, C.POINTER(C.c_void_p), C.c_int32,
                                           C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mbqc_psr_grad_batch_multicast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                                C.c_int64, C.c_void_p, C.c_double, C.c_void_p,
                                                C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mbqc_run_batch_dm_expect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                           C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mbqc_plan_set_feedforward": (C.c_int, [C.c_void_p, C.POINTER(FeedForwardC), C.c_int32]),
    "mbqc_run_batch_sv_sampled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
                                            C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mbqc_run_batch_dm_sampled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
                                            C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mbqc_psr_grad_dataset_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int64]),
    "mbqc_psr_grad_dataset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "mbqc_train_dataset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double,
                                     C.POINTER(Optimizer), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "mbqc_stream_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_int32, C.c_int32,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.c_void_p, C.c_double, C.c_void_p]),
    "mbqc_stream_steps": (C.c_int, [C.c_void_p, C.POINTER(StreamDesc), C.c_void_p]),
    "mbqc_stream_steps_lanes": (C.c_int, [C.c_void_p, C.POINTER(StreamDesc), C.c_void_p]),
    "mbqc_stream_steps_seeded": (C.c_int, [C.c_void_p, C.POINTER(StreamDesc), C.POINTER(StreamSeed), C.c_void_p]),
    "mbqc_stream_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double,
                                       C.c_double, C.c_uint64, C.c_int32, C.c_uint64, C.c_int32,
                                       C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p]),
    "mbqc_stream_gather": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_int32, C.POINTER(C.c_int32),
                                     C.c_void_p, C.c_void_p]),
    "mbqc_device_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "mbqc_device_free": (C.c_int, [C.c_void_p]),
    "mbqc_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mbqc_ipc_import": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mbqc_ipc_close": (C.c_int, [C.c_void_p]),
    "mbqc_peer_barrier": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p]),
    "mbqc_partial_trace_pure": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_void_p]),
    "mbqc_partial_trace_mixed": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_void_p]),
    "mbqc_pure2density": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "mbqc_plan_window": (C.c_int32, [C.c_void_p]),
    "mbqc_plan_num_steps": (C.c_int32, [C.c_void_p]),
    "mbqc_plan_num_outputs": (C.c_int32, [C.c_void_p]),
    "mbqc_probe_fp64_fma": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "mbqc_probe_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "mbqc_launch_count": (C.c_int64, []),
    "mbqc_last_error": (C.c_char_p, []),
    "mbqc_version": (C.c_char_p, []),
}

_lib = None


def load():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension has not been built "
            "(run ./build.sh or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "mentpy_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return list(_SIGNATURES)


def check(rc: int):
    """Map a C status code to the exception the reference would raise for the same misuse."""
    if rc == MBQC_OK:
        return
    msg = load().mbqc_last_error().decode()
    if rc == MBQC_E_ARG:
        raise ValueError(msg)
    if rc == MBQC_E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def launch_count() -> int:
    return int(load().mbqc_launch_count())
