"""Multi-GPU batch parallelism: one process per GPU (torch.distributed, NCCL over NVLink).

Angle sets -- and the +-shift evaluations of a gradient -- are independent pattern runs
(mentpy/gradients/_parameter_shift.py:20-24 has no cross term), so the batch is cut into contiguous
slices, every rank runs the single-GPU kernels on its slice with a replicated plan, and the only
communication is ONE all_gather of the results at the end (or none when the caller keeps results
sharded).  The reference's only parallelism is a process pool over input states in a docs snippet
(docs/tutorials/intro-to-mbqml-parallel.rst:23-51); this is its multi-GPU counterpart.
"""
import os
from typing import Optional, Tuple

import numpy as np


def slice_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first `total % world` ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(total), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _rank_world(group):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_slices(local, total: int, group=None):
    """all_gather of per-rank slices (torch tensors, leading axis = batch) into the full batch on
    every rank.  Ragged slices are padded to the largest one for the collective."""
    import torch
    import torch.distributed as dist

    rank, world = _rank_world(group)
    if world == 1:
        return local
    sizes = [slice_bounds(total, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    if total % world == 0 and local.is_contiguous():
        # even split: the collective writes straight into the result, no padding or re-assembly
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        flat_in = torch.view_as_real(local).reshape(-1) if local.is_complex() else local.reshape(-1)
        flat_out = torch.view_as_real(out).reshape(-1) if out.is_complex() else out.reshape(-1)
        dist.all_gather_into_tensor(flat_out, flat_in, group=group)
        return out
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world, biggest) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    flat_in = torch.view_as_real(pad).reshape(-1) if pad.is_complex() else pad.reshape(-1)
    flat_out = torch.view_as_real(out).reshape(-1) if out.is_complex() else out.reshape(-1)
    dist.all_gather_into_tensor(flat_out, flat_in.contiguous(), group=group)
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def run_batch_distributed(simulator, angles, group=None, gather: bool = True, **kwargs):
    """Evaluate the rows of `angles` ([B,T], identical on every rank) split across the ranks of
    `group`.  Returns the full [B, ...] result on every rank (gather=True, one NCCL all_gather) or
    this rank's slice and its bounds (gather=False)."""
    import torch

    rank, world = _rank_world(group)
    sim = getattr(simulator, "simulator", simulator)
    B = len(angles)
    lo, hi = slice_bounds(B, rank, world)
    dev = sim._dev()
    part = angles[lo:hi]
    if not isinstance(part, torch.Tensor):
        part = torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64))
    local = sim.run_batch(part.to(dev), **kwargs)
    if not gather:
        return local, (lo, hi)
    return gather_slices(local, B, group)


class ReplicatedResult:
    """A [rows, cols] float64 result that every GPU of the node holds a full copy of, filled by
    kernels that store their rows into ALL copies -- the gather is part of the producing launch
    instead of a collective after it.  One process per GPU.  Two transports: an NVSwitch multicast
    address over a torch symmetric-memory allocation (one `multimem.st` stream, replicated by the
    switch), or peer pointers from CUDA IPC over plain cudaMalloc buffers whose handles travel once
    through torch.distributed's object all_gather.  `barrier()` is the stream-ordered flag barrier that
    makes the remote rows visible (mbqc_peer_barrier).  Two copies alternate (`next_copy`), so ONE
    barrier per producing call is enough: a rank can only start overwriting copy c two calls later,
    after a barrier that every rank entered behind its own reads of copy c."""

    _FLAG_BYTES = 256

    def __init__(self, rows: int, cols: int, group=None, device=None, multicast: Optional[bool] = None):
        """multicast: None = use an NVSwitch multicast mapping when torch's symmetric memory provides
        one (every store replicated by the switch: 1/world of the NVLink egress), else peer pointers
        from CUDA IPC; True = require it; False = CUDA IPC only."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        self._lib, self.lib, self.torch = _lib, _lib.load(), torch
        self.group = group
        self.rank, self.world = _rank_world(group)
        if self.world > 8:
            raise ValueError("ReplicatedResult covers one node (<= 8 GPUs)")
        self.rows, self.cols = int(rows), int(cols)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.nbytes = ((self.rows * self.cols * 8 + 255) // 256) * 256
        self._flag_ptrs = None
        self._dst = [None, None]
        self._imported = []
        self.mc_ptr = 0
        self._symm = None
        want_mc = multicast if multicast is not None else os.environ.get("MBQC_MULTICAST", "1") != "0"
        if want_mc and self.world > 1 and self._try_symmetric_memory():
            return
        if multicast:
            raise RuntimeError("no multicast mapping available (torch symmetric memory / NVSwitch multicast)")
        with torch.cuda.device(self.device):
            p = C.c_void_p()
            _lib.check(self.lib.mbqc_device_alloc(2 * self.nbytes + self._FLAG_BYTES, C.byref(p)))
            self._own = p.value
            flags = self.as_tensor(self._own + 2 * self.nbytes, (self._FLAG_BYTES // 8,), torch.int64)
            flags.zero_()
            torch.cuda.synchronize(self.device)
            self.ptrs = [self._own]
            if self.world > 1:
                h = (C.c_char * 64)()
                _lib.check(self.lib.mbqc_ipc_export(C.c_void_p(self._own), h))
                gathered = [None] * self.world
                dist.all_gather_object(gathered, bytes(h), group=group)
                self.ptrs = []
                for r, hb in enumerate(gathered):
                    if r == self.rank:
                        self.ptrs.append(self._own)
                        continue
                    q = C.c_void_p()
                    _lib.check(self.lib.mbqc_ipc_import((C.c_char * 64).from_buffer_copy(hb), C.byref(q)))
                    self._imported.append(q.value)
                    self.ptrs.append(q.value)
                dist.barrier(group=group)  # every flag array is zeroed and mapped before the first use
        self.copy = 0  # the copy the last producing call filled
        self.tensors = [self.as_tensor(self._own + c * self.nbytes, (self.rows, self.cols), torch.float64) for c in (0, 1)]

    def _try_symmetric_memory(self) -> bool:
        """Allocate copies + flags as ONE torch symmetric-memory tensor: peer pointers and, on NVSwitch
        systems, a multicast address come out of its rendezvous.  False (and no side effects) when
        that is not available; the CUDA-IPC path takes over."""
        torch = self.torch
        import torch.distributed as dist

        t = hdl = None
        mc, ptrs = 0, []
        try:
            import torch.distributed._symmetric_memory as symm

            grp = self.group if self.group is not None else dist.group.WORLD
            words = (2 * self.nbytes + self._FLAG_BYTES) // 8
            with torch.cuda.device(self.device):
                t = symm.empty(words, dtype=torch.float64, device=self.device)
                hdl = symm.rendezvous(t, grp)
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            if mc and len(ptrs) == self.world:
                t.zero_()
                torch.cuda.synchronize(self.device)
        except Exception:
            mc = 0
        # every rank takes the same transport: one failure sends all of them to the CUDA-IPC path
        votes = [None] * self.world
        dist.all_gather_object(votes, bool(mc and len(ptrs) == self.world), group=self.group)
        if not all(votes):
            return False
        self._symm, self._symm_hdl = t, hdl
        self.mc_ptr = mc
        self.ptrs = ptrs
        self._own = ptrs[self.rank]
        self.copy = 0
        self.tensors = [self.as_tensor(self._own + c * self.nbytes, (self.rows, self.cols), torch.float64) for c in (0, 1)]
        return True

    @property
    def tensor(self):
        return self.tensors[self.copy]

    def next_copy(self) -> int:
        self.copy ^= 1
        return self.copy

    def as_tensor(self, ptr: int, shape, dtype):
        """torch view of raw device memory (no ownership)."""
        torch = self.torch
        typestr = {torch.float64: "<f8", torch.int64: "<i8"}[dtype]

        class _Raw:
            __cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                        "data": (int(ptr), False), "version": 2, "strides": None}

        return torch.as_tensor(_Raw(), device=self.device)

    def destinations(self):
        """(ctypes array of result pointers with the local copy first, count) for *_push entry points."""
        import ctypes as C

        if self._dst[self.copy] is None:
            off = self.copy * self.nbytes
            # local copy first, then the peers in ring order from this rank: in every phase of the
            # producing kernels the sender -> receiver map is a permutation (no GPU is everyone's target)
            order = [self.ptrs[(self.rank + k) % self.world] + off for k in range(self.world)]
            self._dst[self.copy] = (C.c_void_p * len(order))(*order)
        return self._dst[self.copy], self.world

    def barrier(self, stream=None):
        import ctypes as C

        if self.world == 1:
            return
        torch = self.torch
        if self._flag_ptrs is None:
            self._flag_ptrs = (C.c_void_p * self.world)(*[q + 2 * self.nbytes for q in self.ptrs])
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self._lib.check(self.lib.mbqc_peer_barrier(self._flag_ptrs, self.world, self.rank, st))

    def release(self):
        import ctypes as C

        if getattr(self, "_own", None) is None:
            return
        self.torch.cuda.synchronize(self.device)
        if self._symm is not None:  # torch owns the symmetric allocation
            self._symm, self._symm_hdl = None, None
            self._own, self.tensors = None, [None, None]
            return
        for q in self._imported:
            self.lib.mbqc_ipc_close(C.c_void_p(q))
        self._imported = []
        self.lib.mbqc_device_free(C.c_void_p(self._own))
        self._own, self.tensors = None, [None, None]

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def psr_gradient_distributed(simulator, angles, target, shift: float = 1.5, group=None, gather: bool = True,
                             fused: bool = True, result: Optional["ReplicatedResult"] = None):
    """BASELINE config 4: parameter-shift gradients of B angle vectors split across the GPUs
    (gradients/_parameter_shift.py:9-25 per vector).

    fused=True (default): the gradient kernel of every rank stores its rows straight into all
    ranks' copies of the [B,T] result over NVLink peer memory (`ReplicatedResult`), followed by one
    flag barrier -- no collective call.  Pass `result` to reuse the replicated buffer across calls
    (it is otherwise created, and cached on the simulator, per (B, T)); the returned tensor is a
    view of one of its two alternating copies, valid until the call after the next one.  fused=False: local kernel + one NCCL all_gather."""
    import ctypes as C

    import torch

    from . import _lib
    from .gradients import psr_gradient_batched

    rank, world = _rank_world(group)
    sim = getattr(simulator, "simulator", simulator)
    B = len(angles)
    lo, hi = slice_bounds(B, rank, world)
    part = angles[lo:hi]
    if not isinstance(part, torch.Tensor):
        part = torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64))
    if not gather or not fused or world == 1:
        local = psr_gradient_batched(sim, part.to(sim._dev()), target, shift=shift)
        if not gather:
            return local, (lo, hi)
        return gather_slices(local, B, group)
    dev = sim._dev()
    lib = _lib.load()
    with torch.cuda.device(dev):
        a, _ = sim._stage_angles(part.to(dev), dev)
        batch, T = a.shape
        if result is None:
            cache = sim.__dict__.setdefault("_replicated_results", {})
            result = cache.get((B, T, id(group)))
            if result is None:
                for old in cache.values():
                    old.release()
                cache.clear()
                result = cache[(B, T, id(group))] = ReplicatedResult(B, T, group=group, device=dev)
        if (result.rows, result.cols) != (B, T):
            raise ValueError("result buffer has the wrong shape")
        inp, mode = sim._stage_inputs(None, batch, dev)
        dplan = sim._full_plan()
        tgt = torch.as_tensor(np.ascontiguousarray(target, dtype=np.complex128)).to(dev) \
            if not isinstance(target, torch.Tensor) else target.to(device=dev, dtype=torch.complex128).contiguous()
        status = torch.empty(max(batch, 1), dtype=torch.int32, device=dev)
        result.next_copy()
        st = torch.cuda.current_stream(dev).cuda_stream
        rc = _lib.MBQC_E_UNSUPPORTED
        if result.mc_ptr:  # one store per row, replicated by the switch into every GPU's copy
            rc = lib.mbqc_psr_grad_batch_multicast(dplan.handle, a.data_ptr(), (a.stride(0) if batch > 1 else max(T, 1)),
                                                   None if inp is None else inp.data_ptr(), mode, batch, tgt.data_ptr(),
                                                   C.c_double(shift), C.c_void_p(result.mc_ptr + result.copy * result.nbytes),
                                                   lo, None, status.data_ptr(), st)
            if rc not in (0, _lib.MBQC_E_UNSUPPORTED):
                _lib.check(rc)
        if rc == _lib.MBQC_E_UNSUPPORTED:
            dst, n = result.destinations()
            _lib.check(lib.mbqc_psr_grad_batch_push(dplan.handle, a.data_ptr(), (a.stride(0) if batch > 1 else max(T, 1)),
                                                    None if inp is None else inp.data_ptr(), mode, batch, tgt.data_ptr(),
                                                    C.c_double(shift), dst, n, lo, None, status.data_ptr(), st))
        result.barrier(st)
        sim.last_status = status[:batch]
        return result.tensor


def sample_batch_distributed(simulator, angles, group=None, seed: Optional[int] = None, sample_offset: int = 0,
                             **kwargs):
    """Sampled shots (force0=False) split across the ranks of `group`.  Shot b draws from the
    Philox stream (seed, sample_offset + b) wherever it runs, so the gathered outcome records are
    identical to a single-GPU `sample_batch(angles, seed=seed, sample_offset=sample_offset)`.
    Returns the full SampledBatch on every rank (CUDA tensors)."""
    import torch

    from .simulators.cuda_backends import SampledBatch

    rank, world = _rank_world(group)
    sim = getattr(simulator, "simulator", simulator)
    B = len(angles)
    lo, hi = slice_bounds(B, rank, world)
    part = angles[lo:hi]
    if not isinstance(part, torch.Tensor):
        part = torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64))
    ins = kwargs.pop("input_states", None)
    if ins is not None and getattr(ins, "ndim", 1) == 2:
        ins = ins[lo:hi]
    local = sim.sample_batch(part.to(sim._dev()), input_states=ins, seed=seed, sample_offset=sample_offset + lo, **kwargs)
    return SampledBatch(*[gather_slices(x, B, group) for x in local])


def psr_gradient_dataset_distributed(simulator, angles, targets, input_states=None, shift: float = 1.5, group=None,
                                     return_cost: bool = False):
    """Data-set averaged gradient with the S data items split across the GPUs: every rank runs the
    fused kernel on its items for all P parameter vectors; ONE all_reduce(sum) of P*(T+1) doubles
    combines the partial means ('for a dataset-averaged cost gradient, all-reduce of T
    doubles')."""
    import torch
    import torch.distributed as dist

    from .gradients import psr_gradient_dataset

    rank, world = _rank_world(group)
    sim = getattr(simulator, "simulator", simulator)
    S = len(targets)
    lo, hi = slice_bounds(S, rank, world)
    dev = sim._dev()
    a = angles if isinstance(angles, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(angles, dtype=np.float64))
    a = a.to(dev)
    if input_states is None and getattr(sim, 'input_state', None) is not None and not sim._input_is_plus():
        # differentiate the cost run_batch evaluates (the simulator's own input state)
        input_states = np.tile(np.asarray(sim.input_state, dtype=np.complex128), (S, 1))
    if hi > lo:
        g, c = psr_gradient_dataset(sim, a, targets[lo:hi], None if input_states is None else input_states[lo:hi],
                                    shift=shift, return_cost=True)
        w = (hi - lo) / S
        packed = torch.cat([g.reshape(-1) * w, c.reshape(-1) * w])
    else:
        P = 1 if a.dim() == 1 else a.shape[0]
        packed = torch.zeros(P * (a.shape[-1] + 1), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(packed, group=group)
    n_c = 1 if a.dim() == 1 else a.shape[0]
    g = packed[: packed.numel() - n_c].reshape(a.shape)
    c = packed[packed.numel() - n_c:] if a.dim() == 2 else packed[-1]
    return (g, c) if return_cost else g
