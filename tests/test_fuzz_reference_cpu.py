"""CPU, build container only (needs /root/reference): the oracle against the LIVE reference on
seeded random patterns -- composite circuits, random planes, windows and Haar inputs.  This is
what pins oracle/matrix_free.py beyond the committed golden vectors; the GPU twin
(test_cuda_fuzz.py) then compares the CUDA path with the oracle on the same generator."""
import warnings

import numpy as np
import pytest

from conftest import dm_distance, infidelity_pure
from fuzz_patterns import random_pattern
from oracle import matrix_free
from oracle.pattern_data import PatternData
from oracle.ref_shim import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("seed", range(100))
def test_oracle_equals_live_reference(seed):
    mp = import_reference()
    mixed = seed % 2 == 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            gs, w, ang, inp = random_pattern(mp, seed, mixed)
            ps = mp.PatternSimulator(gs, input_state=inp, backend="numpy-dm" if mixed else "numpy-sv", window_size=w)
        except Exception as e:  # generator hit a combination the reference itself rejects
            pytest.skip(f"reference rejects the pattern: {type(e).__name__}")
        pat = PatternData.from_circuit(gs)
        if mixed:
            want = ps.run(ang)
            got, oc = matrix_free.run_dm_batch(pat, ang[None], input_states=inp[None], window_size=ps.window_size,
                                               return_outcomes=True)
            assert dm_distance(got[0], want) < 1e-10
            assert [int(o) for o in oc[0]] == [int(ps.outcomes[v]) for v in ps.schedule_measure]
        else:
            want = ps.run(ang, output_form="sv")
            got = matrix_free.run_sv_batch(pat, ang[None], input_states=inp[None], window_size=ps.window_size)[0]
            assert infidelity_pure(got, want) < 1e-10
            assert np.allclose(got, want, atol=1e-9)  # including the reference's global phase


@pytest.mark.parametrize("seed", range(100))
def test_host_indexing_equals_live_reference(seed):
    """Same seed through both circuit layers: node labels, inputs/outputs, trainable nodes,
    measurement order and the lowered plan are identical (integer work: bit-exact)."""
    import mentpy_b200 as mb
    from mentpy_b200.plan import lower

    mp = import_reference()
    mixed = seed % 2 == 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            ref, w, ang, _ = random_pattern(mp, seed, mixed)
        except Exception as e:
            pytest.skip(f"reference rejects the pattern: {type(e).__name__}")
        mine, w2, ang2, _ = random_pattern(mb, seed, mixed)
    assert w == w2 and np.array_equal(ang, ang2)
    assert list(ref.graph.nodes()) == list(mine.graph.nodes())
    assert sorted(map(sorted, ref.graph.edges())) == sorted(map(sorted, mine.graph.edges()))
    for attr in ("input_nodes", "output_nodes", "quantum_output_nodes", "trainable_nodes", "measurement_order"):
        assert list(getattr(ref, attr)) == list(getattr(mine, attr)), attr
    a, b = lower(ref, window_size=w, mixed=mixed), lower(mine, window_size=w, mixed=mixed)
    assert (a.schedule, a.input_slot, a.output_slot, a.init_cz_mask) == (b.schedule, b.input_slot, b.output_slot, b.init_cz_mask)
    for sa, sb in zip(a.steps, b.steps):
        assert (sa.node, sa.slot, sa.angle_idx, sa.plane, sa.append, sa.new_node, sa.nbr_mask) == \
               (sb.node, sb.slot, sb.angle_idx, sb.plane, sb.append, sb.new_node, sb.nbr_mask)
        assert sa.fixed_cos == sb.fixed_cos and sa.fixed_sin == sb.fixed_sin


@pytest.mark.parametrize("seed", range(100, 140))
def test_user_schedules_equal_live_reference(seed):
    """`schedule=` kwarg: a perturbed measurement order through the live reference, the oracle and
    the numpy execution of the lowered plan (slot bookkeeping under non-default orders, including
    the CZs the reference silently drops when a neighbour has already left the window)."""
    import mentpy_b200 as mb
    from fuzz_patterns import random_schedule
    from mentpy_b200.plan import lower
    from plan_emulator import run_dm, run_sv

    mp = import_reference()
    mixed = seed % 2 == 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, w, ang, inp = random_pattern(mp, seed, mixed)
        mine, _, _, _ = random_pattern(mb, seed, mixed)
        sched = random_schedule(ref, seed, mixed)
        assert sched == random_schedule(mine, seed, mixed)
        try:
            ps = mp.PatternSimulator(ref, input_state=inp, backend="numpy-dm" if mixed else "numpy-sv",
                                     window_size=w, schedule=sched)
            want = ps.run(ang) if mixed else ps.run(ang, output_form="sv")
        except Exception as e:
            pytest.skip(f"reference rejects the schedule: {type(e).__name__}")
    pat = PatternData.from_circuit(mine)
    pl = lower(mine, window_size=w, schedule=sched, mixed=mixed)
    if mixed:
        got = matrix_free.run_dm_batch(pat, ang[None], input_states=inp[None], window_size=w, schedule=sched)[0]
        assert dm_distance(got, want) < 1e-10
        assert np.abs(run_dm(pl, ang, inp)[0] - want).max() < 1e-10
    else:
        got = matrix_free.run_sv_batch(pat, ang[None], input_states=inp[None], window_size=w, schedule=sched)[0]
        assert infidelity_pure(got, want) < 1e-10
        assert 1 - abs(np.vdot(run_sv(pl, ang, inp), want)) ** 2 < 1e-10


def test_reference_circuits_with_controlled_and_xyz_nodes_lower_like_ours():
    """INTEGRATION.md promises that a mentpy.MBQCircuit works with the CUDA backends as it is: the
    lowering of the REFERENCE's circuit objects (its own ControlMent / MentOutcome / XYZ Ment classes)
    must equal the lowering of this package's mirror, step record by step record."""
    from dataclasses import asdict

    from mentpy_b200.plan import lower
    from oracle.gen_golden import CONTROL_CASES

    mp = import_reference()
    from mentpy.operators import ControlMent as RefControlMent

    import mentpy_b200 as mb

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, build in CONTROL_CASES.items():
            a, b = lower(build(mp, RefControlMent), mixed=True), lower(build(mb, mb.ControlMent), mixed=True)
            assert a.schedule_measure == b.schedule_measure and a.output_slot == b.output_slot, name
            assert [asdict(s) for s in a.steps] == [asdict(s) for s in b.steps], name
        ra, rb = mp.templates.grid_cluster(2, 4), mb.templates.grid_cluster(2, 4)
        ra[2], rb[2] = mp.Ment((0.3, 1.1), "XYZ"), mb.Ment((0.3, 1.1), "XYZ")
        assert [asdict(s) for s in lower(ra, mixed=True).steps] == [asdict(s) for s in lower(rb, mixed=True).steps]
