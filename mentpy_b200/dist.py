"""Multi-GPU batch parallelism: one process per GPU (torch.distributed, NCCL over NVLink).

Angle sets -- and the +-shift evaluations of a gradient -- are independent pattern runs
(mentpy/gradients/_parameter_shift.py:20-24 has no cross term), so the batch is cut into contiguous
slices, every rank runs the single-GPU kernels on its slice with a replicated plan, and the only
communication is ONE all_gather of the results at the end (or none when the caller keeps results
sharded).  The reference's only parallelism is a process pool over input states in a docs snippet
(docs/tutorials/intro-to-mbqml-parallel.rst:23-51); this is its multi-GPU counterpart.
"""
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


def psr_gradient_distributed(simulator, angles, target, shift: float = 1.5, group=None, gather: bool = True):
    """BASELINE config 4: parameter-shift gradients of B angle vectors split across the GPUs."""
    import torch

    from .gradients import psr_gradient_batched

    rank, world = _rank_world(group)
    sim = getattr(simulator, "simulator", simulator)
    B = len(angles)
    lo, hi = slice_bounds(B, rank, world)
    part = angles[lo:hi]
    if not isinstance(part, torch.Tensor):
        part = torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64))
    local = psr_gradient_batched(sim, part.to(sim._dev()), target, shift=shift)
    if not gather:
        return local, (lo, hi)
    return gather_slices(local, B, group)


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
