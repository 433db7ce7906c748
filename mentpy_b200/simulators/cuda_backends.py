"""CUDA backends behind the PatternSimulator facade: `cuda-sv` and `cuda-dm`.

Drop-in twins of NumpySimulatorSV (mentpy/simulators/np_simulator_sv.py:35-384) and
NumpySimulatorDM (np_simulator_dm.py:29-380): same constructor kwargs (`window_size`, `schedule`,
`force0`, `dev_mode`, `wires`; unknown kwargs ignored), same stateful `measure / run / reset`
contract and exceptions, same return conventions (fresh complex128 numpy arrays, big-endian
output order).  All arithmetic runs in the sm_100a kernels through the C ABI
(include/mbqc_b200.h); there is no CPU path.  Additive API: `run_batch` evaluates B angle vectors
in one launch (numpy in -> numpy out with host<->device copies; torch CUDA tensor in -> torch out,
fully asynchronous on the current stream).
"""
import ctypes as C
from collections import namedtuple
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..plan import DevicePlan, LoweredPlan, feedforward, kraus_set, lower, noise_from_kraus
from .backend_base import BaseSimulator

# result of a sampled run: states [B,2^k] / [B,2^k,2^k]; outcomes [B,M] (schedule order); x, z
# byproduct bits [B,k] (output-node order); prob [B] = probability of the outcome record
SampledBatch = namedtuple("SampledBatch", ["states", "outcomes", "x", "z", "prob"])

_NOISE_KINDS = ("depolarizing", "amplitude_damping", "phase_damping", "phase_flip", "bit_flip",
                "generalized_amplitude_damping")


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("mentpy_b200 needs a CUDA device: there is no CPU fallback.")


def _row_stride(a: torch.Tensor) -> int:
    # a size-1 batch axis may report any stride (even 0)
    return a.stride(0) if a.shape[0] > 1 else max(a.shape[1], 1)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class _HostCall:
    """Everything a steady-state host-array `run_batch` call needs, resolved once per
    (plan, batch, output form): rotating sets of (device workspace, page-locked output buffer with
    its numpy view, raw pointers).  The hot call is then: pointer of the input, one ctypes call,
    return a view.  Four sets: up to three calls in flight (run_batch_async) while the caller still
    reads the result of the one before."""

    DEPTH = 4

    def __init__(self, lib, dplan, dev, batch, T, code):
        dim = 2 ** dplan.n_out
        self.shape = (batch, dim) if code == _lib.OUT_SV else (batch, dim, dim)
        out_elems = int(np.prod(self.shape[1:])) * batch
        self.need = int(lib.mbqc_host_workspace_bytes(dplan.handle, batch, code))
        self.d_work = [torch.empty(self.need, dtype=torch.uint8, device=dev) for _ in range(self.DEPTH)]
        self.work_ptr = [t.data_ptr() for t in self.d_work]
        self.h_out = [torch.empty(max(out_elems, 1), dtype=torch.complex128).pin_memory() for _ in range(self.DEPTH)]
        self.out_ptr = [t.data_ptr() for t in self.h_out]
        self.views = [t[:out_elems].numpy().reshape(self.shape) for t in self.h_out]
        self.h_in = [None] * self.DEPTH
        self.flag = C.c_int32(0)
        self.flag_ref = C.byref(self.flag)
        self.ticket = C.c_int32(-1)
        self.ticket_ref = C.byref(self.ticket)
        self.turn = 0

    def staging(self, batch, T):
        if self.h_in[self.turn] is None:
            self.h_in[self.turn] = torch.empty((batch, max(T, 1)), dtype=torch.float64).pin_memory()
        return self.h_in[self.turn]


class PendingBatch:
    """Handle of an asynchronous host call (CudaSimulatorSV.run_batch_async)."""

    def __init__(self, lib, call, turn, ticket, check, keep):
        self._lib, self._call, self._turn, self._ticket, self._check = lib, call, turn, ticket, check
        # what the queued work still reads: the angle buffer (copy engines), the input-state tensor
        # (kernels), and -- through self._call -- the workspace / page-locked result buffers, which
        # therefore outlive an eviction of the simulator's call cache
        self._keep = keep
        self._done = False
        self.status_any = None

    def result(self, copy: bool = False) -> np.ndarray:
        """Block until the call finished and return its output ([B,2^k] or [B,2^k,2^k]).  With
        copy=False this is a view of a rotating page-locked buffer, valid until three further calls
        have been submitted."""
        if not self._done:
            flag = C.c_int32(0)
            _lib.check(self._lib.mbqc_host_wait(self._ticket, C.byref(flag)))
            self._done, self._keep, self.status_any = True, None, flag.value
            if self._check and (flag.value & _lib.STATUS_BAD_NORM):
                raise ValueError("qstate has nan, you might want to increase the window size")
        res = self._call.views[self._turn]
        return res.copy() if copy else res

    def __del__(self):
        # a handle dropped without result() must still give its ticket back (four tickets per device)
        if not getattr(self, "_done", True):
            try:
                self._lib.mbqc_host_wait(self._ticket, None)
            except Exception:
                pass
            self._done = True


class _ReadyBatch:
    def __init__(self, value):
        self._value, self.status_any = value, 0

    def result(self, copy: bool = False):
        return self._value


class _CudaPatternBase(BaseSimulator):
    mixed = False

    def __init__(self, mbqcircuit, input_state: np.ndarray = None, **kwargs) -> None:
        super().__init__(mbqcircuit, input_state)
        self.window_size = kwargs.pop("window_size", 1)
        self.schedule = kwargs.pop("schedule", None)
        self.force0 = kwargs.pop("force0", True)
        self.dev_mode = kwargs.pop("dev_mode", False)
        self.wires = kwargs.pop("wires", None)
        self.device = kwargs.pop("device", None)
        self._slot_order = kwargs.pop("slot_order", "msb")
        dtype = kwargs.pop("dtype", "complex128")
        dtype = {"complex128": "complex128", "complex64": "complex64", np.complex128: "complex128",
                 np.complex64: "complex64"}.get(dtype, str(dtype))
        if dtype not in ("complex128", "complex64"):
            raise ValueError("dtype must be 'complex128' or 'complex64'")
        self.dtype = dtype
        # force0=False (NotImplementedError in the reference, np_simulator_sv.py:50-51) samples the
        # outcomes and applies the flow corrections: see sample_batch
        self.seed = int(kwargs.pop("seed", 0))
        self._shots_done = 0
        self._noise = self._parse_noise(kwargs)
        # dev_mode (np_simulator_sv.py:173-203, np_simulator_dm.py:160-201): the window picks the measured node
        # by the wire rule; it only changes the order of the steps, which the lowering works out up front
        self.plan: LoweredPlan = lower(mbqcircuit, self.window_size, self.schedule, mixed=self.mixed,
                                       slot_order=self._slot_order, dev_mode=bool(self.dev_mode), wires=self.wires)
        self.window_size = self.plan.window
        self.schedule = self.plan.schedule
        self.schedule_measure = self.plan.schedule_measure
        if not self.force0:  # fail early (no GPU needed): flow, planes and window of the sampled path
            wmax = _lib.MAX_WINDOW_REG if self.mixed else _lib.MAX_WINDOW_SMEM_SV
            if self.plan.window > wmax:
                raise NotImplementedError(f"force0=False covers window_size <= {wmax}")
            if any(st.cond_mask for st in self.plan.steps):
                raise NotImplementedError("force0=False does not cover outcome-controlled measurements (ControlMent).")
            feedforward(mbqcircuit, self.plan)
        if input_state is None:
            n_in = len(mbqcircuit.input_nodes)
            input_state = np.full(2**n_in, 2.0 ** (-n_in / 2))
        self.input_state = np.asarray(input_state)
        self.current_measurement = 0
        self._angles_seen = np.zeros(max(self.plan.n_angles, 1))
        self._full: Optional[DevicePlan] = None
        self._prefix = {}
        self._d_input = None
        self.last_status = None
        self._host_calls = {}
        self._input_synced = False

    # -- plumbing -------------------------------------------------------------------------------
    def _parse_noise(self, kwargs):
        return None

    def _dev(self) -> torch.device:
        _require_cuda()
        if self.device is None:
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device(self.device)

    def _full_plan(self) -> DevicePlan:
        if self._full is None:
            with torch.cuda.device(self._dev()):
                self._full = DevicePlan(self.plan, self._noise)
        return self._full

    def _prefix_plan(self, n_done: int) -> Tuple[DevicePlan, List[int]]:
        """Plan of the first n_done measurements whose 'outputs' are the whole live window in the
        reference's order (position 0 = MSB) -- backs `measure()` / `qstate`."""
        if n_done not in self._prefix:
            nodes = self.plan.window_nodes_after(n_done)
            slot = self.plan.slot_of_after(n_done)
            with torch.cuda.device(self._dev()):
                self._prefix[n_done] = (DevicePlan(self.plan, self._noise_for_prefix(), n_steps=n_done,
                                                   output_slot=[slot[v] for v in nodes]), nodes)
        return self._prefix[n_done]

    def _noise_for_prefix(self):
        return None

    def _input_is_plus(self) -> bool:
        """True when the simulator's input state is the default |+>^|I| (pattern_simulator.py:58-61)."""
        n = len(self.plan.input_nodes)
        st = np.asarray(self.input_state)
        return st.shape == (2 ** n,) and np.allclose(st, np.full(2 ** n, 2.0 ** (-n / 2)), rtol=0, atol=1e-15)

    def _device_input(self, dev):
        st = np.ascontiguousarray(self.input_state, dtype=np.complex128)
        if st.shape != (2 ** len(self.plan.input_nodes),):
            raise ValueError(
                f"Input state has shape {st.shape}, expected ({2 ** len(self.plan.input_nodes)},)."
            )
        return torch.from_numpy(st).to(dev)

    def _stage_inputs(self, input_states, batch, dev):
        """-> (tensor|None, input_mode)"""
        if input_states is None:
            if self._d_input is None or self._d_input.device != dev:
                self._d_input = self._device_input(dev)
                # the default |+>^|I| input needs no loads at all: the kernels seed constants
                self._plus_input = self._input_is_plus()
            if self._plus_input:
                return None, _lib.INPUT_PLUS
            return self._d_input, _lib.INPUT_SHARED
        if isinstance(input_states, torch.Tensor):
            t = input_states.to(device=dev, dtype=torch.complex128).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(input_states, dtype=np.complex128)).to(dev)
        dim = 2 ** len(self.plan.input_nodes)
        if t.dim() == 1:
            if t.shape[0] != dim:
                raise ValueError(f"input state must have {dim} amplitudes")
            return t, _lib.INPUT_SHARED
        if t.shape != (batch, dim):
            raise ValueError(f"input_states must have shape ({batch}, {dim}) or ({dim},)")
        return t, _lib.INPUT_BATCH

    def _stage_angles(self, angles, dev):
        if isinstance(angles, torch.Tensor):
            a = angles.to(device=dev, dtype=torch.float64)
            if a.dim() == 1:
                a = a[None, :]
            if a.stride(-1) != 1:
                a = a.contiguous()
            on_host = False
        else:
            a = np.asarray(angles, dtype=np.float64)
            if a.ndim == 1:
                a = a[None, :]
            a = torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
            on_host = True
        if a.dim() != 2 or a.shape[1] != self.plan.n_angles:
            raise ValueError(
                f"Number of angles ({a.shape[-1]}) does not match number of trainable nodes ({self.plan.n_angles})."
            )
        return a, on_host

    def _check_status(self, status: torch.Tensor):
        st = status.cpu().numpy()
        self.last_status = st
        if (st & _lib.STATUS_BAD_NORM).any():
            raise ValueError("qstate has nan, you might want to increase the window size")
        return st

    # -- sampled runs (force0=False) ---------------------------------------------------------------
    def _sampling_plan(self) -> DevicePlan:
        dplan = self._full_plan()
        if not dplan.has_feedforward:
            dplan.set_feedforward(feedforward(self.mbqcircuit, self.plan))
        return dplan

    def sample_batch(self, angles, input_states=None, output_form: Optional[str] = None, seed: Optional[int] = None,
                     sample_offset: Optional[int] = None, forced_outcomes=None, correct: bool = True):
        """Run B shots with Born-rule outcomes (or the given `forced_outcomes` [B,M]) and flow
        corrections; see csrc/sample.cuh.  Returns a `SampledBatch` (states, outcomes [B,M] in
        schedule order, x / z byproduct bits [B,k] per output node, probability of each record);
        numpy in -> numpy out, CUDA tensors in -> CUDA tensors out.

        Shot b uses the Philox stream (seed, sample_offset + b); by default sample_offset continues
        where the previous call stopped, so repeated calls give fresh, reproducible shots.  With
        correct=True the byproducts are applied to the outputs (noiseless shots then all equal the
        force0 state up to a global phase); with correct=False the raw branch state is returned."""
        dev = self._dev()
        lib = _lib.load()
        if self.dtype != "complex128":
            raise NotImplementedError("sampled runs are complex128 only")
        wmax = _lib.MAX_WINDOW_REG if self.mixed else _lib.MAX_WINDOW_SMEM_SV
        if self.plan.window > wmax:
            raise NotImplementedError(f"sampled runs cover window_size <= {wmax}")
        form = (output_form or ("dm" if self.mixed else "sv")).lower()
        if form not in ("sv", "statevector", "dm", "densitymatrix"):
            raise ValueError(f"Output form {output_form} is not supported.")
        want_dm = form in ("dm", "densitymatrix")
        if self.mixed and not want_dm:
            raise ValueError("the density-matrix backend returns density matrices")
        with torch.cuda.device(dev):
            dplan = self._sampling_plan()
            a, on_host = self._stage_angles(angles, dev)
            batch, M, k = a.shape[0], dplan.n_steps, dplan.n_out
            inp, mode = self._stage_inputs(input_states, batch, dev)
            dim = 2 ** k
            out = torch.empty((batch, dim, dim) if self.mixed else (batch, dim), dtype=torch.complex128, device=dev)
            if forced_outcomes is None:
                outc = torch.zeros((batch, max(M, 1)), dtype=torch.int8, device=dev)
                omode = _lib.OUTCOMES_SAMPLE
            else:
                outc = torch.as_tensor(forced_outcomes).to(device=dev, dtype=torch.int8).reshape(batch, -1).contiguous()
                if outc.shape[1] != M:
                    raise ValueError(f"forced_outcomes must have shape ({batch}, {M})")
                if M == 0:
                    outc = torch.zeros((batch, 1), dtype=torch.int8, device=dev)
                omode = _lib.OUTCOMES_FORCED
            byp = torch.zeros(batch, dtype=torch.int32, device=dev)
            prob = torch.empty(batch, dtype=torch.float64, device=dev)
            status = torch.empty(batch, dtype=torch.int32, device=dev)
            if sample_offset is None:
                sample_offset = self._shots_done
                self._shots_done += batch
            fn = lib.mbqc_run_batch_dm_sampled if self.mixed else lib.mbqc_run_batch_sv_sampled
            _lib.check(fn(dplan.handle, _ptr(a), _row_stride(a), _ptr(inp), mode, batch,
                          C.c_uint64(self.seed if seed is None else int(seed)), C.c_uint64(int(sample_offset)),
                          omode, 1 if correct else 0, _ptr(out), _ptr(outc), _ptr(byp), _ptr(prob), _ptr(status),
                          torch.cuda.current_stream(dev).cuda_stream))
            if want_dm and not self.mixed:
                out = out[:, :, None] * out.conj()[:, None, :]
            q = torch.arange(k, device=dev, dtype=torch.int32)
            bx = ((byp[:, None] >> q[None, :]) & 1).to(torch.int8)
            bz = ((byp[:, None] >> (q[None, :] + 16)) & 1).to(torch.int8)
            outc = outc[:, :M]
            if on_host:
                self._check_status(status)
                return SampledBatch(out.cpu().numpy(), outc.cpu().numpy(), bx.cpu().numpy(), bz.cpu().numpy(),
                                    prob.cpu().numpy())
            self.last_status = status
            return SampledBatch(out, outc, bx, bz, prob)

    def _sampled_run(self, angles, output_form):
        res = self.sample_batch(np.asarray(angles, dtype=np.float64)[None, :], output_form=output_form)
        self.outcomes = {v: int(o) for v, o in zip(self.schedule_measure, res.outcomes[0])}
        self.byproducts = {v: (int(x), int(z)) for v, x, z in zip(self.plan.output_nodes, res.x[0], res.z[0])}
        return res.states[0]

    def outcome_averaged(self, angles, input_state=None, max_measurements: int = 20) -> np.ndarray:
        """sum over ALL outcome records b of p_b rho_b (corrected outputs): the state the reference's
        PennyLane backend returns, where measurements are deferred and corrections are gates
        (pennylane_simulator.py:113-153) -- with a channel that is the noisy full-graph result.
        Every record is forced through the sampled kernels (2^M shots, 65,536 per launch); the
        branch probabilities sum to 1 (checked).  Density-matrix form [2^k, 2^k]."""
        M = len(self.plan.steps)
        if M > max_measurements:
            raise NotImplementedError(f"2^{M} outcome records (max_measurements = {max_measurements})")
        ang = np.asarray(angles, dtype=np.float64).reshape(1, -1)
        k = len(self.plan.output_nodes)
        acc, total = np.zeros((2**k, 2**k), dtype=complex), 0.0
        chunk = 1 << 16
        for lo in range(0, 1 << M, chunk):
            ids = np.arange(lo, min(lo + chunk, 1 << M), dtype=np.int64)
            recs = ((ids[:, None] >> (M - 1 - np.arange(M))[None, :]) & 1).astype(np.int8).reshape(len(ids), M)
            res = self.sample_batch(np.repeat(ang, len(ids), 0), input_states=input_state, forced_outcomes=recs,
                                    output_form="dm", correct=True)
            acc += np.einsum("b,bij->ij", res.prob, res.states)
            total += float(res.prob.sum())
        if abs(total - 1.0) > 1e-9:
            raise ValueError(f"branch probabilities sum to {total}")
        return acc

    # -- reference-compatible state machine -----------------------------------------------------
    def reset(self, input_state: np.ndarray = None):
        self.current_measurement = 0
        if input_state is not None:
            self.input_state = np.asarray(input_state)
            self._d_input = None
            self._input_synced = False
        self._angles_seen[:] = 0.0
        self.outcomes = {}
        self._shots_done += 1  # a stateful run is one shot of the Philox stream (plane-Z draws, mode="sample")

    def current_simulated_nodes(self) -> List[int]:
        return self.plan.window_nodes_after(self.current_measurement)

    def current_number_simulated_nodes(self) -> int:
        return min(self.window_size, len(self.mbqcircuit) - self.current_measurement)

    def _record_angle(self, angle):
        if self.current_measurement >= len(self.schedule_measure):
            raise ValueError("No more measurements to be done.")
        st = self.plan.steps[self.current_measurement]
        if st.plane == _lib.PLANE_Z:
            return st  # angle-free
        if st.cond_mask:  # controlled node: its angle column, whichever branch ends up using it
            if angle is None:
                raise ValueError("Measurement is trainable, please provide an angle.")
            self._angles_seen[st.column] = float(angle)
        elif st.angle_idx >= 0:
            if angle is None:
                raise ValueError("Measurement is trainable, please provide an angle.")
            self._angles_seen[st.angle_idx] = float(angle)
        elif angle is not None and angle != st.fixed_angle:
            raise ValueError(f"Measurement has a fixed angle of {round(st.fixed_angle, 4)}")
        return st

    def find_swaps(self, source, target):
        """Selection-sort swap list source -> target (np_simulator_sv.py:360-374)."""
        assert set(source) == set(target), (
            f"Both lists must have the same elements, but source={source} and target={target}"
        )
        work, swaps = list(source), []
        for i, want in enumerate(target):
            if work[i] != want:
                j = work.index(want, i + 1)
                work[i], work[j] = work[j], work[i]
                swaps.append((i, j))
        return swaps


class CudaSimulatorSV(_CudaPatternBase):
    """State-vector backend (`backend="cuda-sv"`)."""

    mixed = False

    def run_batch(self, angles, input_states=None, output_form: str = "sv", check: bool = True,
                  copy: bool = True):
        """Evaluate B angle vectors: angles [B,T] -> [B,2^k] ('sv') or [B,2^k,2^k] ('dm').

        Host input (numpy / CPU tensor, ideally pinned -- see `mentpy_b200.pinned_empty`) returns a
        numpy array; with copy=False it is a view of an internal pinned buffer that stays valid
        until the next-but-one host call.  CUDA tensor input returns a CUDA tensor, asynchronously
        on the current stream."""
        form = output_form.lower()
        if form in ("dm", "densitymatrix"):
            code = _lib.OUT_DM
        elif form in ("sv", "statevector"):
            code = _lib.OUT_SV
        else:
            raise ValueError(f"Output form {output_form} is not supported.")
        return self._run_plan(self._full_plan(), angles, input_states, code, check, copy)

    def _run_plan(self, dplan: DevicePlan, angles, input_states, code, check, copy=True):
        dev = self._dev()
        lib = _lib.load()
        if self.dtype == "complex64":
            return self._run_plan_f32(dplan, angles, input_states, code, check)
        on_host = not (isinstance(angles, torch.Tensor) and angles.is_cuda)
        if on_host and (input_states is None or np.ndim(input_states) == 1):
            return self._run_plan_host(dplan, angles, input_states, code, check, copy)
        with torch.cuda.device(dev):
            a, on_host = self._stage_angles(angles, dev)
            batch = a.shape[0]
            inp, mode = self._stage_inputs(input_states, batch, dev)
            dim = 2 ** dplan.n_out
            shape = (batch, dim) if code == _lib.OUT_SV else (batch, dim, dim)
            out = torch.empty(shape, dtype=torch.complex128, device=dev)
            status = torch.empty(batch, dtype=torch.int32, device=dev)
            _lib.check(lib.mbqc_run_batch_sv(dplan.handle, _ptr(a), _row_stride(a), _ptr(inp), mode,
                                             batch, _ptr(out), code, _ptr(status),
                                             torch.cuda.current_stream(dev).cuda_stream))
            if on_host:
                res = out.cpu().numpy()
                if check:
                    self._check_status(status)
                return res
            self.last_status = status
            return out

    def _run_plan_f32(self, dplan, angles, input_states, code, check):
        """complex64 mode (mbqc_run_batch_sv_f32): fp32 state, complex64 outputs."""
        dev = self._dev()
        lib = _lib.load()
        if self.plan.window > _lib.MAX_WINDOW_REG:
            raise NotImplementedError(f"complex64 mode covers window <= {_lib.MAX_WINDOW_REG}")
        with torch.cuda.device(dev):
            a, on_host = self._stage_angles(angles, dev)
            batch = a.shape[0]
            inp, mode = self._stage_inputs(input_states, batch, dev)
            inp = None if inp is None else inp.to(torch.complex64).contiguous()
            dim = 2 ** dplan.n_out
            shape = (batch, dim) if code == _lib.OUT_SV else (batch, dim, dim)
            out = torch.empty(shape, dtype=torch.complex64, device=dev)
            status = torch.empty(max(batch, 1), dtype=torch.int32, device=dev)
            if batch:
                _lib.check(lib.mbqc_run_batch_sv_f32(dplan.handle, _ptr(a), _row_stride(a), _ptr(inp), mode,
                                                     batch, _ptr(out), code, _ptr(status),
                                                     torch.cuda.current_stream(dev).cuda_stream))
            if on_host:
                res = out.cpu().numpy()
                if check and batch:
                    self._check_status(status[:batch])
                return res
            self.last_status = status[:batch]
            return out

    def _run_plan_host(self, dplan, angles, input_states, code, check, copy, submit_only=False):
        """Host arrays in, host arrays out through the C-level chunked H2D/kernel/D2H pipeline
        (mbqc_run_batch_sv_host)."""
        dev = self._dev()
        lib = _lib.load()
        if isinstance(angles, torch.Tensor):
            src = angles if angles.dtype == torch.float64 else angles.to(torch.float64)
        else:
            src = torch.from_numpy(np.ascontiguousarray(np.asarray(angles, dtype=np.float64)))
        if src.dim() == 1:
            src = src[None, :]
        if src.dim() != 2 or src.shape[1] != self.plan.n_angles:
            raise ValueError(
                f"Number of angles ({src.shape[-1]}) does not match number of trainable nodes ({self.plan.n_angles})."
            )
        if not src.is_contiguous():
            src = src.contiguous()
        batch, T = src.shape
        if batch == 0:
            dim = 2 ** dplan.n_out
            empty = np.zeros((0, dim) if code == _lib.OUT_SV else (0, dim, dim), dtype=np.complex128)
            return _ReadyBatch(empty) if submit_only else empty
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        key = (id(dplan), batch, code)
        call = self._host_calls.get(key)
        if call is None:
            if len(self._host_calls) > 8:
                # in-flight PendingBatch handles hold their own reference to their _HostCall, so the
                # buffers of queued work stay alive; only idle entries are really dropped here
                self._host_calls.clear()
            call = self._host_calls[key] = _HostCall(lib, dplan, dev, batch, T, code)
        inp, mode = self._stage_inputs(input_states, batch, dev)
        if input_states is not None or not self._input_synced:
            # the pipeline reads `inp` on its own non-blocking streams: a tensor that was just
            # uploaded on torch's current stream must be complete before the submit (a per-call
            # input state is a fresh temporary every time; the cached default only once)
            torch.cuda.current_stream(dev).synchronize()
            self._input_synced = True
        call.turn = (call.turn + 1) % call.DEPTH
        if batch * T >= (1 << 16) and not src.is_pinned():
            stage = call.staging(batch, T)  # large pageable input: one copy into page-locked memory
            stage.copy_(src)
            src = stage
        if submit_only:
            rc = lib.mbqc_run_batch_sv_host_submit(dplan.handle, src.data_ptr(), max(T, 1), _ptr(inp), mode, batch,
                                                   call.out_ptr[call.turn], code, call.work_ptr[call.turn], call.need,
                                                   0, call.ticket_ref)
            if rc:
                _lib.check(rc)
            self.last_status = None
            return PendingBatch(lib, call, call.turn, call.ticket.value, check, (src, inp))
        rc = lib.mbqc_run_batch_sv_host(dplan.handle, src.data_ptr(), max(T, 1), _ptr(inp), mode, batch,
                                        call.out_ptr[call.turn], code, call.work_ptr[call.turn], call.need,
                                        call.flag_ref, 0)
        if rc:
            _lib.check(rc)
        if check and (call.flag.value & _lib.STATUS_BAD_NORM):
            raise ValueError("qstate has nan, you might want to increase the window size")
        self.last_status = None
        res = call.views[call.turn]
        return res.copy() if copy else res

    def run_batch_async(self, angles, input_states=None, output_form: str = "sv", check: bool = True) -> "PendingBatch":
        """Queue a host-array `run_batch` and return at once; `.result()` of the returned handle
        blocks until the output is complete.  Up to three calls may be in flight: the host-to-device
        copy of call n+1 then overlaps the kernels and result transfer of call n (PCIe is full
        duplex), which roughly halves the per-call time of a stream of batches.  `angles` must be
        a host array (numpy / CPU tensor, ideally page-locked) and must stay untouched until
        `.result()`; complex128 only."""
        form = output_form.lower()
        if form in ("dm", "densitymatrix"):
            code = _lib.OUT_DM
        elif form in ("sv", "statevector"):
            code = _lib.OUT_SV
        else:
            raise ValueError(f"Output form {output_form} is not supported.")
        if self.dtype != "complex128":
            raise NotImplementedError("run_batch_async is complex128 only")
        if isinstance(angles, torch.Tensor) and angles.is_cuda:
            raise ValueError("run_batch_async takes host arrays; CUDA tensors are already asynchronous in run_batch")
        return self._run_plan_host(self._full_plan(), angles, input_states, code, check, False, submit_only=True)

    def measure(self, angle: float) -> Tuple[np.ndarray, int]:
        if not self.force0:
            raise NotImplementedError("step-by-step measure() is deterministic (force0=True) only; use run / sample_batch")
        st = self._record_angle(angle)
        self.current_measurement += 1
        self.outcomes[st.node] = 0
        return self.qstate, 0

    @property
    def qstate(self) -> np.ndarray:
        """Window state in the reference's layout after `current_measurement` measurements."""
        dplan, _nodes = self._prefix_plan(self.current_measurement)
        return self._run_plan(dplan, self._angles_seen[None, : self.plan.n_angles], None, _lib.OUT_SV, True)[0]

    def run(self, angles: List[float], output_form="dm", **kwargs):
        if kwargs.get("input_state") is not None:
            self.reset(input_state=kwargs.get("input_state"))
        if len(angles) != len(self.mbqcircuit.trainable_nodes):
            raise ValueError(
                f"Number of angles ({len(angles)}) does not match number of trainable nodes ({len(self.mbqcircuit.trainable_nodes)})."
            )
        if self.current_measurement != 0:
            raise ValueError("No more measurements to be done.")
        if not self.force0:
            res = self._sampled_run(angles, output_form)
            self.current_measurement = len(self.schedule_measure)
            return res
        res = self.run_batch(np.asarray(angles, dtype=np.float64)[None, :], output_form=output_form)[0]
        self.current_measurement = len(self.schedule_measure)
        self._angles_seen[: self.plan.n_angles] = np.asarray(angles, dtype=np.float64)
        self.outcomes = {v: 0 for v in self.schedule_measure}
        return res

    def reorder_qubits(self, state, current_order, target_order):
        """Permute qubits of a host state vector (np_simulator_sv.py:376-384)."""
        n = len(current_order)
        t = np.asarray(state).reshape([2] * n)
        return t.transpose([list(current_order).index(v) for v in target_order]).reshape(-1)


class CudaSimulatorDM(_CudaPatternBase):
    """Density-matrix backend (`backend="cuda-dm"`), optional single-qubit noise
    (`circuit_noise=<kind>`, `p=`, `gamma=`, `p_gad=` or `kraus=[...]`)."""

    mixed = True

    def _parse_noise(self, kwargs):
        kind = kwargs.pop("circuit_noise", None)
        kraus = kwargs.pop("kraus", None)
        p = kwargs.pop("p", 0.0)
        gamma = kwargs.pop("gamma", None)
        p_gad = kwargs.pop("p_gad", 0.5)
        self.circuit_noise = kind
        if kraus is not None:
            return noise_from_kraus(kraus)
        if kind is None:
            return None
        if kind not in _NOISE_KINDS:
            raise ValueError(f"Unrecognized circuit noise: {kind}")
        return noise_from_kraus(kraus_set(kind, p=p, gamma=gamma, p_gad=p_gad))

    def _noise_for_prefix(self):
        return self._noise

    def run_batch(self, angles, input_states=None, check: bool = True, return_outcomes: bool = False,
                  mode: str = "sample", seed: Optional[int] = None, sample_offset: Optional[int] = None):
        """angles [B,T] -> rho [B,2^k,2^k] (and the outcome record [B,M] if requested).

        Patterns with plane-Z nodes (np_simulator_dm.py:327-346): in mode="expectation" those qubits
        are traced out unprojected and their entry of the outcome record is prob1 (the record is
        then float64); in mode="sample" their outcome is drawn from (prob0, prob1), even under
        force0, and the state projected -- row b draws from the Philox stream (seed, sample_offset +
        b), by default continuing where the previous call stopped (parity with the reference's
        np.random draws is statistical only)."""
        return self._run_plan(self._full_plan(), angles, input_states, check, return_outcomes, mode,
                              seed=seed, sample_offset=sample_offset)

    def _has_z(self, n_steps):
        return any(st.plane == _lib.PLANE_Z for st in self.plan.steps[:n_steps])

    def _run_plan(self, dplan, angles, input_states, check, return_outcomes, zmode="sample", seed=None,
                  sample_offset=None, advance=True):
        dev = self._dev()
        lib = _lib.load()
        has_z = self._has_z(dplan.n_steps)
        expect = has_z and zmode in ("expectation", "exp")
        zsample = has_z and not expect
        with torch.cuda.device(dev):
            a, on_host = self._stage_angles(angles, dev)
            batch = a.shape[0]
            inp, mode = self._stage_inputs(input_states, batch, dev)
            dim = 2 ** dplan.n_out
            out = torch.empty((batch, dim, dim), dtype=torch.complex128, device=dev)
            status = torch.empty(batch, dtype=torch.int32, device=dev)
            # every kernel writes the whole record (no memset launch in front of a ~10 us kernel)
            outc = torch.empty((batch, max(dplan.n_steps, 1)), dtype=torch.int8, device=dev)
            if expect:
                zp = torch.empty((batch, max(dplan.n_steps, 1)), dtype=torch.float64, device=dev)
                _lib.check(lib.mbqc_run_batch_dm_expect(dplan.handle, _ptr(a), _row_stride(a), _ptr(inp), mode,
                                                        batch, _ptr(out), _ptr(outc), _ptr(zp), _ptr(status),
                                                        torch.cuda.current_stream(dev).cuda_stream))
                outc = outc.to(torch.float64) + zp  # plane-Z entries: prob1; the others stay 0 / 1
            elif zsample:
                if sample_offset is None:
                    sample_offset = self._shots_done
                    if advance:
                        self._shots_done += batch
                _lib.check(lib.mbqc_run_batch_dm_zsample(dplan.handle, _ptr(a), _row_stride(a), _ptr(inp), mode, batch,
                                                         C.c_uint64(self.seed if seed is None else int(seed)),
                                                         C.c_uint64(int(sample_offset)), _ptr(out), _ptr(outc), _ptr(status),
                                                         torch.cuda.current_stream(dev).cuda_stream))
            else:
                _lib.check(lib.mbqc_run_batch_dm(dplan.handle, _ptr(a), _row_stride(a), _ptr(inp), mode,
                                                 batch, _ptr(out), _ptr(outc), _ptr(status),
                                                 torch.cuda.current_stream(dev).cuda_stream))
            if on_host:
                res = out.cpu().numpy()
                oc = outc.cpu().numpy()[:, : dplan.n_steps]
                if check:
                    self._check_status(status)
                return (res, oc) if return_outcomes else res
            self.last_status = status
            return (out, outc[:, : dplan.n_steps]) if return_outcomes else out

    def measure(self, angle: float, mode="sample") -> Tuple[np.ndarray, int]:
        st = self._record_angle(angle)
        self.current_measurement += 1
        self._last_mode = mode
        dplan, _nodes = self._prefix_plan(self.current_measurement)
        # the prefix is re-run from the seed state: the plane-Z draws of earlier steps must repeat, so
        # the whole run uses ONE shot index (advanced by reset())
        rho, oc = self._run_plan(dplan, self._angles_seen[None, : self.plan.n_angles], None, True, True, mode,
                                 sample_offset=self._shots_done, advance=False)
        outcome = oc[0, self.current_measurement - 1]
        outcome = float(outcome) if st.plane == _lib.PLANE_Z else int(outcome)
        self.outcomes[st.node] = outcome
        return rho[0], outcome

    @property
    def qstate(self) -> np.ndarray:
        dplan, _nodes = self._prefix_plan(self.current_measurement)
        return self._run_plan(dplan, self._angles_seen[None, : self.plan.n_angles], None, True, False,
                              getattr(self, "_last_mode", "expectation"), sample_offset=self._shots_done, advance=False)[0]

    def run(self, angles: List[float], mode="sample", input_state=None):
        if input_state is not None:
            self.reset(input_state=input_state)
        if len(angles) != len(self.mbqcircuit.trainable_nodes):
            raise ValueError(
                f"Number of angles ({len(angles)}) does not match number of trainable nodes ({len(self.mbqcircuit.trainable_nodes)})."
            )
        if self.current_measurement != 0:
            raise ValueError("No more measurements to be done.")
        if not self.force0:
            res = self._sampled_run(angles, "dm")
            self.current_measurement = len(self.schedule_measure)
            return res
        rho, oc = self.run_batch(np.asarray(angles, dtype=np.float64)[None, :], return_outcomes=True, mode=mode)
        self.current_measurement = len(self.schedule_measure)
        self._angles_seen[: self.plan.n_angles] = np.asarray(angles, dtype=np.float64)
        self.outcomes = {st.node: (float(o) if st.plane == _lib.PLANE_Z else int(o))
                         for st, o in zip(self.plan.steps, oc[0])}
        return rho[0]

    def reorder_qubits(self, state, current_order, target_order):
        n = len(current_order)
        perm = [list(current_order).index(v) for v in target_order]
        st = np.asarray(state)
        if st.ndim == 1:
            return st.reshape([2] * n).transpose(perm).reshape(-1)
        t = st.reshape([2] * (2 * n))
        return t.transpose(perm + [n + q for q in perm]).reshape(2**n, 2**n)
