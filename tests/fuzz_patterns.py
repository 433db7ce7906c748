"""Seeded random MBQC patterns built through the PUBLIC circuit API (templates, hstack, vstack,
merge, Ment), usable with the reference package (`mp`) and with mentpy_b200 (`mb`) alike -- the
generator only calls names both expose.  Used to fuzz the oracle against the live reference (CPU,
build container) and the CUDA path against the oracle (GPU)."""
import numpy as np


def random_pattern(lib, seed: int, mixed: bool):
    """-> (circuit, window_size, angles [T], input_state [2^|I|])."""
    from scipy.stats import unitary_group

    rng = np.random.default_rng(seed)
    T = lib.templates

    def base():
        kind = rng.integers(0, 7)
        if kind == 0:
            return T.linear_cluster(int(rng.integers(3, 8)))
        if kind == 1:
            return T.grid_cluster(2, int(rng.integers(2, 6)))
        if kind == 2:
            return T.grid_cluster(3, int(rng.integers(2, 5)))
        if kind == 3:
            return T.muta(2, 1, one_column=True)
        if kind == 4:
            return T.many_wires([int(x) for x in rng.integers(2, 5, size=int(rng.integers(1, 4)))])
        if kind == 5:
            return T.grid_cluster(int(rng.integers(2, 4)), int(rng.integers(3, 5)), periodic=True)
        return T.muta(2, 1)

    gs = base()
    if len(gs.measurement_order) - len(gs.output_nodes) <= len(gs.input_nodes) + 1:
        gs = T.linear_cluster(5)  # too few measurements for any window: both simulators refuse those
    op = rng.integers(0, 4)
    if op == 1:  # extend every wire with another block of the same height
        rows = len(gs.output_nodes)
        tail = T.linear_cluster(int(rng.integers(2, 5))) if rows == 1 else T.grid_cluster(rows, int(rng.integers(2, 4)))
        gs = lib.hstack([gs, tail])
    elif op == 2 and len(gs.input_nodes) <= 2:  # independent extra wire
        gs = lib.vstack([gs, T.linear_cluster(int(rng.integers(2, 5)))])
    elif op == 3 and len(gs.output_nodes) >= 1:  # glue a wire onto the first output
        other = T.linear_cluster(int(rng.integers(2, 5)))
        gs = lib.merge(gs, other, along=[(gs.output_nodes[0], other.input_nodes[0])])
    nodes = [v for v in gs.measurement_order if v not in gs.output_nodes]
    picks = rng.permutation(nodes)[: min(3, len(nodes))]
    planes = ["X", "Y", "fixed"] + (["XZ", "YZ"] if mixed else [])
    for v in picks:
        if len(gs.trainable_nodes) <= 2:
            break
        choice = planes[int(rng.integers(0, len(planes)))]
        if choice == "fixed":
            gs[int(v)] = lib.Ment(float(rng.uniform(0, 2 * np.pi)), "XY")
        elif choice in ("XZ", "YZ"):
            gs[int(v)] = lib.Ment(float(rng.uniform(0, 2 * np.pi)), choice) if rng.integers(0, 2) else lib.Ment(choice)
        else:
            gs[int(v)] = lib.Ment(choice)
    if mixed:
        # every third density-matrix pattern also gets a fixed two-angle XYZ node (ment.py:239-251); drawn from
        # its own stream so that the patterns of the other seeds stay what they were
        r2 = np.random.default_rng(20_000 + seed)
        free = [int(v) for v in gs.trainable_nodes if v not in gs.output_nodes]
        if r2.integers(0, 3) == 0 and len(free) > 2:
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                gs[free[int(r2.integers(0, len(free)))]] = lib.Ment((float(r2.uniform(0, 2 * np.pi)), float(r2.uniform(-1.5, 1.5))), "XYZ")
    n_in = len(gs.input_nodes)
    n_meas = len(gs.measurement_order) - len(gs.quantum_output_nodes if mixed else gs.output_nodes)
    hi = min(n_meas, 6 if mixed else 8)   # register, shared-memory and (SV) streaming kernels
    lo = min(n_in + 1, hi)
    window = max(int(rng.integers(lo, hi + 1)), 2)  # window_size=1 means "default" in the reference
    angles = rng.uniform(0, 2 * np.pi, len(gs.trainable_nodes))
    inp = unitary_group.rvs(2**n_in, random_state=int(seed))[:, 0] if n_in else np.ones(1, dtype=complex)
    if seed % 5 == 0:  # the default |+> input (pattern_simulator.py:58-61)
        inp = np.full(2**n_in, 2.0 ** (-n_in / 2), dtype=complex)
    return gs, window, angles, inp


def random_schedule(gs, seed: int, mixed: bool):
    """A user schedule (the `schedule=` kwarg of the simulators): the default measurement order with
    a few random adjacent transpositions among the measured non-input nodes; inputs stay first and
    the unmeasured outputs last, as both simulators require."""
    rng = np.random.default_rng(10_000 + seed)
    order = list(gs.measurement_order)
    keep_last = list(gs.quantum_output_nodes if mixed else gs.output_nodes)
    n_in = len(gs.input_nodes)
    head, tail = order[:n_in], [v for v in order[n_in:] if v in keep_last]
    mid = [v for v in order[n_in:] if v not in keep_last]
    for _ in range(int(rng.integers(1, 4))):
        if len(mid) >= 2:
            i = int(rng.integers(0, len(mid) - 1))
            mid[i], mid[i + 1] = mid[i + 1], mid[i]
    return head + mid + tail
