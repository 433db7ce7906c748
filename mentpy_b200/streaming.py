"""Streaming / sharded state-vector execution of one large-window pattern (`cuda-sv-stream`).

Host half of the streaming regime (include/mbqc_b200.h "streaming regime"): turns the lowered
plan (mentpy_b200/plan.py) into a list of passes -- fused local passes and NVLink exchange steps --
and drives an engine that owns the device buffers.  One process per GPU; with G = 2^g ranks the top
g bits of the 2^w index are the rank.  The reference cannot run these windows at all (it builds
2^w x 2^w operators, mentpy/operators/gates.py:62-72,127-143), so parity for this path is pinned by
the analytic linear-cluster oracle, the matrix-free numpy oracle at small w, and 1-GPU == sharded
agreement (tests/test_streaming_host.py, tests/test_cuda_streaming.py).

Physical layout per rank: the local share (2^L amplitudes, L = w - g) is kept as two half
buffers H0 / H1 selected by the top local bit, plus one spare half when sharded.  A measurement
whose slot is a shard bit is executed as ONE kernel per rank that reads the partner's half over
NVLink and keeps both results local; this moves the appended qubit to the top local slot and the
qubit that lived there to the shard slot, so the executor keeps a logical->physical slot map.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from .plan import LoweredPlan

MAX_FUSE = 5
MAX_RANGES = 16
LANE_BITS = 5   # index bits that map to lanes of a warp


@dataclass
class LocalPass:
    """K fused measurements on local slots: one read + one write of the live local amplitudes."""
    slots: List[int]                # physical slots, in measurement order
    cos_t: List[float]
    sin_t: List[float]
    nbr_masks: List[int]            # physical, whole index
    local_masks: List[int]          # restricted to the fused slots (bit i = measurement i)
    append_mask: int
    ranges: List[tuple]             # (pos, width) zero fields, ascending
    n_groups: int
    scale: float
    steps: List[int]                # plan step indices (bookkeeping)
    live_bits: int = 0              # live local bits before the pass (traffic accounting)
    dead_bits: List[int] = field(default_factory=list)   # dead local slots before the pass
    lane: bool = False              # all fused slots < 5, no dead slot < 5: warp-shuffle kernel


@dataclass
class ExchangePass:
    """One measurement on a shard slot, fused with the peer read."""
    shard_bit: int                  # which rank bit pairs the partners
    cos_t: float
    sin_t: float
    append: bool
    half_mask: int                  # neighbour slots among the half-index bits
    rank_mask: int                  # neighbour slots among the rank bits (parity from the rank)
    top_is_neighbour: bool          # neighbour mask contains the top local slot
    scale: float
    step: int
    dead_bits: List[int] = field(default_factory=list)   # dead local slots before the pass


@dataclass
class StreamSchedule:
    window: int
    local_bits: int
    shard_bits: int
    passes: List[object]
    output_slots: List[int]         # physical slots of the output qubits, first = MSB
    dead_shard_bits: List[int]      # rank bits whose slot died: only ranks with 0 there hold data
    phase: complex                  # prod (1 + e^{i theta}) / |.|  (reference global phase)
    algorithmic_bytes: int          # sum over measurements of 2 * 16 * 2^n(live)
    streamed_bytes: int             # bytes the passes actually read + write per job (all ranks)


def _ranges(bits: Sequence[int]) -> List[tuple]:
    out = []
    for b in sorted(bits):
        if out and out[-1][0] + out[-1][1] == b:
            out[-1] = (out[-1][0], out[-1][1] + 1)
        else:
            out.append((b, 1))
    return out


def _step_cos_sin(st, angles):
    if st.angle_idx >= 0:
        th = float(angles[st.angle_idx])
        return float(np.cos(th)), float(np.sin(th))
    return st.fixed_cos, st.fixed_sin


def build_schedule(plan: LoweredPlan, angles, shard_bits: int = 0, fuse: int = 4) -> StreamSchedule:
    """Pure host logic (no GPU): passes for every rank are identical except for the rank-dependent
    constants the engine fills in (index_or, exchange role)."""
    if plan.mixed:
        raise NotImplementedError("the streaming regime covers the state-vector path")
    w = plan.window
    L = w - shard_bits
    if shard_bits < 0 or L < 2:
        raise ValueError(f"window {w} too small for {1 << shard_bits} shards")
    if len(angles) != plan.n_angles:
        raise ValueError(
            f"Number of angles ({len(angles)}) does not match number of trainable nodes ({plan.n_angles})."
        )
    fuse = max(1, min(int(fuse), MAX_FUSE))
    phys: Dict[int, int] = {s: s for s in range(w)}   # logical slot -> physical slot
    dead: set = set()                                 # dead physical slots
    passes: List[object] = []
    appended = 0
    phase = 1.0 + 0.0j
    algo = 0
    streamed = 0
    live = w

    def pmask(logical_mask: int) -> int:
        m, out = logical_mask, 0
        while m:
            b = (m & -m).bit_length() - 1
            out |= 1 << phys[b]
            m &= m - 1
        return out

    def scale_for(n_append: int) -> float:
        nonlocal appended
        k = (appended + n_append) // 2 - appended // 2
        appended += n_append
        return 0.5 ** k

    steps = plan.steps
    i = 0
    while i < len(steps):
        st = steps[i]
        c, s = _step_cos_sin(st, angles)
        ps = phys[st.slot]
        if ps >= L:                                    # ---- shard slot: exchange pass
            z = complex(1.0 + c, s)
            phase *= z / abs(z)
            algo += 2 * 16 * (1 << live)
            pm = pmask(st.nbr_mask) if st.append else 0
            top = L - 1
            passes.append(ExchangePass(
                shard_bit=ps - L, cos_t=c, sin_t=s, append=st.append,
                half_mask=pm & ((1 << top) - 1), rank_mask=pm >> L,
                top_is_neighbour=bool((pm >> top) & 1), scale=scale_for(1 if st.append else 0), step=i,
                dead_bits=sorted(d for d in dead if d < L)))
            live_local = L - len([d for d in dead if d < L])
            active = 1 << (shard_bits - len([d for d in dead if d >= L]))
            if st.append:
                if top in dead:
                    raise NotImplementedError("append onto a shard slot while the top local slot is dead")
                other = next(lg for lg, p in phys.items() if p == top)
                phys[st.slot], phys[other] = top, ps
                # per rank: read own half + peer half (NVLink), write own half + spare half
                streamed += active * 16 * 4 * (1 << (live_local - 1))
            else:
                dead.add(ps)
                live -= 1
                # survivors (half of the active ranks): read own + peer share, write own share
                streamed += (active // 2) * 16 * 3 * (1 << live_local)
            i += 1
            continue
        # ---- local slots: fuse consecutive measurements
        grp, j = [], i
        while j < len(steps) and len(grp) < fuse:
            pj = phys[steps[j].slot]
            if pj >= L or pj in [g[1] for g in grp]:
                break
            if grp and (pj < LANE_BITS) != (grp[0][1] < LANE_BITS):
                break                 # lane-bit slots and register slots go to different kernels
            grp.append((j, pj))   # distinct slots <=> all measured qubits were live at pass start
            j += 1
        slots = [p for _, p in grp]
        cos_l, sin_l, nbr_l, loc_l, amask, n_app = [], [], [], [], 0, 0
        live_before = live
        for k, (sj, pj) in enumerate(grp):
            sst = steps[sj]
            cj, sjn = _step_cos_sin(sst, angles)
            z = complex(1.0 + cj, sjn)
            phase *= z / abs(z)
            algo += 2 * 16 * (1 << live)
            cos_l.append(cj)
            sin_l.append(sjn)
            pm = pmask(sst.nbr_mask) if sst.append else 0
            nbr_l.append(pm)
            loc_l.append(sum(1 << q for q, pq in enumerate(slots) if (pm >> pq) & 1))
            if sst.append:
                amask |= 1 << k
                n_app += 1
            else:
                live -= 1
        dead_local = [d for d in dead if d < L]
        fixed = sorted(set(slots) | set(dead_local))
        rng = _ranges(fixed)
        if len(rng) > MAX_RANGES:
            raise NotImplementedError("too many disjoint fixed-bit ranges for one pass")
        live_local_before = L - len(dead_local)
        n_groups = 1 << (L - len(fixed))
        n_dead_shard = len([d for d in dead if d >= L])
        active_ranks = 1 << (shard_bits - n_dead_shard)
        # reads 2^live_local, writes the survivors
        n_written = (1 << (live_local_before - (len(slots) - n_app)))
        streamed += active_ranks * 16 * ((1 << live_local_before) + n_written)
        lane = (all(p < LANE_BITS for p in slots) and all(dd >= LANE_BITS for dd in dead_local)
                and live_local_before >= LANE_BITS + 2 and L - 1 >= LANE_BITS)
        passes.append(LocalPass(slots, cos_l, sin_l, nbr_l, loc_l, amask, rng, n_groups,
                                scale_for(n_app), [sj for sj, _ in grp], live_local_before,
                                sorted(dead_local), lane))
        for sj, pj in grp:
            if not steps[sj].append:
                dead.add(pj)
        i = j
    out_phys = []
    lg_out = {v: s for v, s in zip(plan.output_nodes, plan.output_slot)}
    for v in plan.output_nodes:
        out_phys.append(phys[lg_out[v]])
    return StreamSchedule(w, L, shard_bits, passes, out_phys, sorted(d - L for d in dead if d >= L),
                          phase, algo, streamed)


class StreamExecutor:
    """Runs a StreamSchedule on an engine (the CUDA engine below, or the numpy engines the host
    tests define).  Engine contract: init(plan, scale), local_pass(p, index_or), exchange(p, role,
    partner, const_parity), gather(output_slots) -> complex[2^k] with only locally-owned entries,
    allreduce(vec), barrier()."""

    def __init__(self, plan: LoweredPlan, engine, rank: int = 0, shard_bits: int = 0, fuse: int = 4):
        self.plan, self.engine, self.rank, self.shard_bits, self.fuse = plan, engine, rank, shard_bits, fuse

    def run(self, angles, input_state=None) -> np.ndarray:
        sched = build_schedule(self.plan, angles, self.shard_bits, self.fuse)
        self.last_schedule = sched
        L, rank = sched.local_bits, self.rank
        eng = self.engine
        # when the pattern starts with a local pass, the engine may generate the seed inside that
        # pass instead of writing it out first (one write + one read of the whole state saved)
        first_local = bool(sched.passes) and isinstance(sched.passes[0], LocalPass)
        import os
        import time

        prof = os.environ.get("MBQC_STREAM_PROFILE") == "1"
        self.timeline = []

        def mark(label, t0):
            if prof:
                eng.barrier()
                self.timeline.append((label, (time.perf_counter() - t0) * 1e3))

        t0 = time.perf_counter()
        eng.init(self.plan, L, rank, input_state, defer=first_local)
        mark("init", t0)
        if first_local:  # the first pass generates its input: no read of the state in that pass
            p0 = sched.passes[0]
            sched.streamed_bytes -= (1 << self.shard_bits) * 16 * (1 << p0.live_bits)
        alive = True
        for n_pass, p in enumerate(sched.passes):
            t0 = time.perf_counter()
            if isinstance(p, LocalPass):
                if alive:
                    eng.local_pass(p, rank << L, seeded=(n_pass == 0 and first_local))
                mark(f"local K={len(p.slots)} slots={p.slots} live={p.live_bits} lane={p.lane}", t0)
                continue
            v = (rank >> p.shard_bit) & 1
            partner = rank ^ (1 << p.shard_bit)
            par = bin(rank & p.rank_mask).count("1") & 1
            if p.top_is_neighbour:
                par ^= v                      # the old top-local qubit has value v on this rank
            eng.barrier()                     # partner's earlier passes are complete and visible
            if p.append:
                if alive:
                    eng.exchange(p, role=v, partner=partner, const_parity=par)
                eng.rotate_roles(p.shard_bit)  # every rank tracks every rank's buffer roles
            else:
                if alive and v == 0:
                    eng.exchange(p, role=2, partner=partner, const_parity=0)
                elif alive and hasattr(eng, "serve_dying"):
                    eng.serve_dying(p, partner)   # emulation engines only: peers cannot read our memory
                if v == 1:
                    alive = False             # this rank's share died with the slot
            eng.barrier()
            mark(f"exchange append={p.append}", t0)
        t0 = time.perf_counter()
        vec = eng.gather(sched.output_slots, L, rank, alive)
        vec = eng.allreduce(vec)
        mark("gather+allreduce", t0)
        nrm = np.linalg.norm(vec)
        if not np.isfinite(nrm) or nrm == 0.0:
            raise ValueError("qstate has nan, you might want to increase the window size")
        return vec / nrm * sched.phase


# ---------------------------------------------------------------------------------------------
# CUDA engine
# ---------------------------------------------------------------------------------------------
class CudaStreamEngine:
    """Owns the half buffers of one rank (plain cudaMalloc through the C ABI so they can be mapped
    by neighbours with CUDA IPC) and launches the streaming kernels."""

    def __init__(self, device=None, group=None):
        import torch

        from . import _lib

        self.torch, self._lib = torch, _lib
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.group = group
        self.bufs: List[int] = []
        self.half_elems = 0
        self.peer_ptrs: Dict[int, List[int]] = {}
        self.roles = {}

    # -- memory ---------------------------------------------------------------------------------
    def _alloc(self, L: int, world: int):
        import ctypes as C

        half = 1 << (L - 1)
        n_buf = 3 if world > 1 else 2
        if self.half_elems == half and len(self.bufs) == n_buf:
            return
        self.release()
        self.half_elems = half
        with self.torch.cuda.device(self.device):
            if world == 1:
                # one contiguous allocation: H1 directly follows H0
                p = C.c_void_p()
                self._lib.check(self.lib.mbqc_device_alloc(2 * half * 16, C.byref(p)))
                self.bufs = [p.value, p.value + half * 16]
                self._owned = [p.value]
            else:
                self.bufs, self._owned = [], []
                for _ in range(3):
                    p = C.c_void_p()
                    self._lib.check(self.lib.mbqc_device_alloc(half * 16, C.byref(p)))
                    self.bufs.append(p.value)
                    self._owned.append(p.value)
        if world > 1:
            self._exchange_handles(world)

    def _exchange_handles(self, world: int):
        import ctypes as C
        import torch.distributed as dist

        handles = []
        for b in self.bufs:
            h = (C.c_char * 64)()
            self._lib.check(self.lib.mbqc_ipc_export(C.c_void_p(b), h))
            handles.append(bytes(h))
        gathered = [None] * world
        dist.all_gather_object(gathered, handles, group=self.group)
        rank = dist.get_rank(self.group)
        self.peer_ptrs = {}
        with self.torch.cuda.device(self.device):
            for r, hs in enumerate(gathered):
                if r == rank:
                    continue
                ptrs = []
                for hb in hs:
                    p = C.c_void_p()
                    buf = (C.c_char * 64).from_buffer_copy(hb)
                    self._lib.check(self.lib.mbqc_ipc_import(buf, C.byref(p)))
                    ptrs.append(p.value)
                self.peer_ptrs[r] = ptrs

    def release(self):
        import ctypes as C

        for ptrs in self.peer_ptrs.values():
            for p in ptrs:
                self.lib.mbqc_ipc_close(C.c_void_p(p))
        self.peer_ptrs = {}
        for p in getattr(self, "_owned", []):
            self.lib.mbqc_device_free(C.c_void_p(p))
        self._owned, self.bufs, self.half_elems = [], [], 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # -- engine contract ------------------------------------------------------------------------
    def _world(self):
        import torch.distributed as dist

        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized() and self.group is not False) else 1

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def init(self, plan: LoweredPlan, L: int, rank: int, input_state, defer: bool = False):
        import ctypes as C

        world = 1 << (plan.window - L)
        self._alloc(L, world)
        # roles[r] = [index of H0, index of H1, index of spare] in rank r's buffer list
        self.roles = {r: [0, 1, 2] for r in range(world)}
        self.L = L
        w = plan.window
        n_in = len(plan.input_slot)
        d_in = None
        if input_state is not None:
            st = np.ascontiguousarray(input_state, dtype=np.complex128)
            if st.shape != (1 << n_in,):
                raise ValueError(f"Input state has shape {st.shape}, expected ({1 << n_in},).")
            d_in = self.torch.from_numpy(st).to(self.device)
            scale = 2.0 ** (-(w - n_in) / 2)
        else:
            scale = 2.0 ** (-w / 2)
        in_arr = (C.c_int32 * max(n_in, 1))(*plan.input_slot)
        cz_arr = (C.c_uint64 * w)(*plan.init_cz_mask)
        seed = self._lib.StreamSeed()
        seed.window, seed.n_inputs, seed.scale = w, n_in, scale
        for q, sl in enumerate(plan.input_slot):
            seed.input_slot[q] = sl
        for a, m in enumerate(plan.init_cz_mask):
            seed.init_cz_mask[a] = m
        seed.d_input = None if d_in is None else d_in.data_ptr()
        self._seed = seed
        self._keep = d_in
        self.rank = rank
        if defer:
            return
        with self.torch.cuda.device(self.device):
            for h in (0, 1):
                index_or = (rank << L) | (h << (L - 1))
                self._lib.check(self.lib.mbqc_stream_init(
                    C.c_void_p(self.bufs[h]), L - 1, index_or, w, n_in, in_arr, cz_arr,
                    None if d_in is None else C.c_void_p(d_in.data_ptr()), scale, C.c_void_p(self._stream())))
        self._keep = d_in
        self.rank = rank

    def _hi_offset(self) -> int:
        r = self.roles[self.rank]
        return ((self.bufs[r[1]] - self.bufs[r[0]]) // 16) & 0xFFFFFFFFFFFFFFFF

    def local_pass(self, p: LocalPass, index_or: int, seeded: bool = False):
        import ctypes as C

        def launch(buf, desc):
            if seeded:
                self._lib.check(self.lib.mbqc_stream_steps_seeded(C.c_void_p(buf), C.byref(desc), C.byref(self._seed),
                                                                  C.c_void_p(self._stream())))
            else:
                self._lib.check(self.lib.mbqc_stream_steps(C.c_void_p(buf), C.byref(desc), C.c_void_p(self._stream())))

        if p.lane and not seeded:
            return self._lane_pass(p, index_or)
        d = self._lib.StreamDesc()
        top = self.L - 1
        hi = self._hi_offset()
        d.n_fused = len(p.slots)
        rng = list(p.ranges)
        # the top local bit selects the half buffer: it is always squeezed out of the thread
        # index; when it is not a fused slot the pass is issued once per half
        top_fused = top in p.slots
        top_dead = any(pos <= top < pos + wd for pos, wd in rng) and not top_fused
        for k, sl in enumerate(p.slots):
            d.elem_offset[k] = hi if sl == top else (1 << sl)
            d.elem_bit[k] = 1 << sl
            d.cos_t[k], d.sin_t[k] = p.cos_t[k], p.sin_t[k]
            d.nbr_mask[k], d.local_mask[k] = p.nbr_masks[k], p.local_masks[k]
        d.append_mask = p.append_mask
        d.scale = p.scale
        r = self.roles[self.rank]
        with self.torch.cuda.device(self.device):
            if top_fused or top_dead:
                d.n_ranges = len(rng)
                for q, (pos, wd) in enumerate(rng):
                    d.range_pos[q], d.range_width[q] = pos, wd
                d.n_groups = p.n_groups
                d.index_or = index_or
                launch(self.bufs[r[0]], d)
            else:
                # ranges never reach the top bit here: run the pass on each half separately
                d.n_ranges = len(rng)
                for q, (pos, wd) in enumerate(rng):
                    d.range_pos[q], d.range_width[q] = pos, wd
                d.n_groups = p.n_groups >> 1
                for h in (0, 1):
                    d.index_or = index_or | (h << top)
                    launch(self.bufs[r[h]], d)

    def _lane_pass(self, p: LocalPass, index_or: int):
        """Fused slots are lane bits: one element per thread, partners via warp shuffles."""
        import ctypes as C

        d = self._lib.StreamDesc()
        top = self.L - 1
        d.n_fused = len(p.slots)
        for k, sl in enumerate(p.slots):
            d.elem_bit[k] = 1 << sl
            d.elem_offset[k] = 1 << sl
            d.cos_t[k], d.sin_t[k] = p.cos_t[k], p.sin_t[k]
            d.nbr_mask[k], d.local_mask[k] = p.nbr_masks[k], p.local_masks[k]
        d.append_mask = p.append_mask
        d.scale = p.scale
        rng = _ranges(p.dead_bits)
        d.n_ranges = len(rng)
        for q, (pos, wd) in enumerate(rng):
            d.range_pos[q], d.range_width[q] = pos, wd
        top_dead = top in p.dead_bits
        n_elems = 1 << (self.L - len(p.dead_bits))
        r = self.roles[self.rank]
        with self.torch.cuda.device(self.device):
            if top_dead:
                d.n_groups, d.index_or = n_elems, index_or
                self._lib.check(self.lib.mbqc_stream_steps_lanes(C.c_void_p(self.bufs[r[0]]), C.byref(d), C.c_void_p(self._stream())))
            else:
                d.n_groups = n_elems >> 1
                for h in (0, 1):
                    d.index_or = index_or | (h << top)
                    self._lib.check(self.lib.mbqc_stream_steps_lanes(C.c_void_p(self.bufs[r[h]]), C.byref(d), C.c_void_p(self._stream())))

    def exchange(self, p: ExchangePass, role: int, partner: int, const_parity: int):
        import ctypes as C

        mine, theirs = self.roles[self.rank], self.roles[partner]
        top = self.L - 1
        dead_half = [d for d in p.dead_bits if d < top]          # dead slots inside the half index
        rng = _ranges(dead_half)
        pos = (C.c_uint32 * max(len(rng), 1))(*[r[0] for r in rng])
        wid = (C.c_uint32 * max(len(rng), 1))(*[r[1] for r in rng])
        n = 1 << (top - len(dead_half))                          # live elements per half
        with self.torch.cuda.device(self.device):
            st = C.c_void_p(self._stream())
            if role == 2:
                for h in ((0,) if top in p.dead_bits else (0, 1)):
                    self._lib.check(self.lib.mbqc_stream_exchange(
                        C.c_void_p(self.bufs[mine[h]]), C.c_void_p(self.peer_ptrs[partner][theirs[h]]), None, 2,
                        p.cos_t, p.sin_t, p.scale, 0, 0, n, len(rng), pos, wid, st))
                return
            own = self.bufs[mine[role]]
            peer = self.peer_ptrs[partner][theirs[role]]
            spare = self.bufs[mine[2]]
            self._lib.check(self.lib.mbqc_stream_exchange(
                C.c_void_p(own), C.c_void_p(peer), C.c_void_p(spare), role, p.cos_t, p.sin_t, p.scale,
                p.half_mask, const_parity, n, len(rng), pos, wid, st))

    def rotate_roles(self, shard_bit: int):
        """After an append exchange on `shard_bit`: ranks with 0 there adopted the spare as H1,
        ranks with 1 adopted it as H0 (old half becomes the new spare)."""
        bit = 1 << shard_bit
        for r, ro in self.roles.items():
            if r & bit:
                ro[0], ro[2] = ro[2], ro[0]
            else:
                ro[1], ro[2] = ro[2], ro[1]

    def gather(self, output_slots, L, rank, alive) -> np.ndarray:
        import ctypes as C

        k = len(output_slots)
        out = self.torch.zeros(1 << k, dtype=self.torch.complex128, device=self.device)
        if alive:
            top = L - 1
            r = self.roles[rank]
            arr = (C.c_int32 * max(k, 1))(*output_slots)
            with self.torch.cuda.device(self.device):
                for h in (0, 1):
                    index_or = (rank << L) | (h << top)
                    self._lib.check(self.lib.mbqc_stream_gather(
                        C.c_void_p(self.bufs[r[h]]), top, index_or, k, arr, C.c_void_p(out.data_ptr()),
                        C.c_void_p(self._stream())))
        self._out = out
        return out

    def allreduce(self, vec):
        import torch.distributed as dist

        if self._world() > 1:
            dist.all_reduce(vec, group=self.group)
        return vec.cpu().numpy()

    def barrier(self):
        import torch.distributed as dist

        self.torch.cuda.synchronize(self.device)
        if self._world() > 1:
            dist.barrier(group=self.group)


class CudaSimulatorSVStream:
    """`backend="cuda-sv-stream"`: one angle set, window up to ~33 qubits per GPU, optionally
    sharded over the ranks of a torch.distributed group (one process per GPU, NCCL)."""

    def __init__(self, mbqcircuit, input_state: np.ndarray = None, **kwargs):
        from .plan import lower

        self.mbqcircuit = mbqcircuit
        self.group = kwargs.get("group", None)
        self.window_size = kwargs.pop("window_size", 1)
        self.schedule = kwargs.pop("schedule", None)
        self.fuse = kwargs.pop("fuse", 5)
        self.group = kwargs.pop("group", None)
        self.force0 = kwargs.pop("force0", True)
        if not self.force0:
            raise NotImplementedError("Numpy simulator does not support force0=False.")
        if kwargs.pop("dev_mode", False):
            # the streaming passes fuse measurements by slot; dev_mode orders are served by cuda-sv / cuda-dm
            raise NotImplementedError("dev_mode scheduling is not supported by the streaming backend.")
        # Sharded runs keep the window positions measured LAST in the shard slots: a shard slot costs
        # an NVLink exchange every time it is measured, so the first w - g measurements are all
        # local (and start on the high, coalesced local slots); for patterns not much longer than
        # the window, shard slots are only reached in the tail when the live state is already tiny.
        world = self._dist()[1]
        slot_order = kwargs.pop("slot_order", None)
        if slot_order is None:
            slot_order = "shard-last" if world > 1 else "msb"
        self.slot_order = slot_order
        self.plan = lower(mbqcircuit, self.window_size, self.schedule, mixed=False, slot_order=slot_order,
                          shard_bits=max(world.bit_length() - 1, 0))
        self.window_size = self.plan.window
        self.schedule = self.plan.schedule
        self.schedule_measure = self.plan.schedule_measure
        self.input_state = input_state
        self.outcomes = {}
        self._engine = None

    def _dist(self):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and self.group is not False:
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def reset(self, input_state=None):
        if input_state is not None:
            self.input_state = input_state
        self.outcomes = {}

    def run(self, angles, output_form="sv", **kwargs):
        if kwargs.get("input_state") is not None:
            self.reset(kwargs["input_state"])
        rank, world = self._dist()
        g = world.bit_length() - 1
        if (1 << g) != world:
            raise ValueError("the number of ranks must be a power of two")
        if self._engine is None:
            self._engine = CudaStreamEngine(group=self.group)
        default_in = self.input_state is None or np.allclose(
            self.input_state, np.full(1 << len(self.plan.input_slot), 2.0 ** (-len(self.plan.input_slot) / 2)))
        ex = StreamExecutor(self.plan, self._engine, rank, g, self.fuse)
        psi = ex.run(np.asarray(angles, dtype=np.float64), None if default_in else self.input_state)
        self.last_schedule = ex.last_schedule
        self.outcomes = {v: 0 for v in self.schedule_measure}
        form = output_form.lower()
        if form in ("dm", "densitymatrix"):
            return np.outer(psi, np.conj(psi))
        if form in ("sv", "statevector"):
            return psi
        raise ValueError(f"Output form {output_form} is not supported.")

    def __call__(self, angles, **kwargs):
        return self.run(angles, **kwargs)
