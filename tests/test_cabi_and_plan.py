"""CPU: the C-ABI library loads and exports every symbol the header declares (no compute calls),
and the host-side plan lowering reproduces the reference's window bookkeeping."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mentpy_b200 as mb
from mentpy_b200 import _lib
from mentpy_b200.plan import kraus_set, lower, noise_from_kraus, window_is_valid
from conftest import ROOT, build_spec, load_golden
from oracle.pattern_data import PatternData


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mbqc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mbqc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mbqc_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared
    assert b"sm_100a" in lib.mbqc_version()


def test_plan_create_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call."""
    import ctypes as C

    lib = _lib.load()
    handle = C.c_void_p()
    step = (_lib.Step * 1)()
    step[0].slot = 5  # outside a 2-qubit window
    cz = (C.c_uint64 * 2)(0, 0)
    io = (C.c_int32 * 1)(0)
    rc = lib.mbqc_plan_create(step, 1, 2, 1, 1, 1, io, cz, io, None, C.byref(handle))
    assert rc == _lib.MBQC_E_ARG and b"slot" in lib.mbqc_last_error()
    rc = lib.mbqc_plan_create(step, 1, 99, 1, 1, 1, io, cz, io, None, C.byref(handle))
    assert rc == _lib.MBQC_E_ARG
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_lowering_matches_reference_window_bookkeeping():
    """Replay the slot plan on the host and compare with the reference's shifting window
    (np_simulator_sv.py:130-142, :207-223) recorded in the golden patterns."""
    for rec in load_golden("structures.json")["records"]:
        gs = build_spec(rec["spec"])
        pat = PatternData.from_json(rec["pattern"])
        if pat.measurement_order is None:   # no causal flow (spturb): nothing to lower without a user schedule
            with pytest.raises(ValueError, match="Schedule must be provided"):
                lower(gs)
            continue
        n_meas = len(pat.measurement_order) - len(pat.output_nodes)
        for w in {len(pat.input_nodes) + 1, min(len(pat.input_nodes) + 3, n_meas)}:
            if w > n_meas or w < len(pat.input_nodes) or w == 1:
                continue  # w == 1 is promoted to |I|+1 (np_simulator_sv.py:74-75)
            pl = lower(gs, window_size=w)
            sched = pat.measurement_order
            assert pl.schedule == sched
            for m, st in enumerate(pl.steps):
                assert st.node == sched[m]
                win_after = sched[m + 1: m + 1 + w]
                assert st.append == (m + 1 + w <= pat.n_nodes)
                slots = pl.slot_of_after(m + 1)
                assert sorted(slots) == sorted(win_after)
                assert len(set(slots.values())) == len(slots)
                if st.append:
                    assert st.new_node == win_after[-1]
                    want = 0
                    for nb in pat.neighbors(st.new_node):
                        if nb in win_after[:-1]:
                            want |= 1 << slots[nb]
                    assert st.nbr_mask == want
            # np_simulator_sv.py:286-290: window order kept when it equals quantum_output_nodes
            tail = sched[len(pl.steps):]
            assert pl.output_nodes == (pat.quantum_output_nodes if pat.quantum_output_nodes == tail else pat.output_nodes)
            assert [pl.slot_of_after(len(pl.steps))[v] for v in pl.output_nodes] == pl.output_slot


def test_lowering_errors_mirror_reference():
    gs = mb.templates.grid_cluster(2, 4)
    with pytest.raises(ValueError, match="Input state has 2 qubits"):
        lower(gs, window_size=1, schedule=gs.measurement_order)
    with pytest.raises(ValueError, match="schedule only has"):
        lower(gs, window_size=7)
    gs[1] = mb.Ment(0.2, "XZ")
    with pytest.raises(ValueError, match="only XY plane is supported"):
        lower(gs)
    lower(gs, mixed=True)
    gs[1] = mb.Ment("Z")  # DM path only (mode="expectation"), angle-free, register kernels only
    zp = lower(gs, mixed=True)
    zs = next(st for st in zp.steps if st.node == 1)
    assert zs.plane == _lib.PLANE_Z and zs.angle_idx == -1 and zp.n_angles == len(gs.trainable_nodes) == 5
    with pytest.raises(ValueError, match="only XY plane is supported"):
        lower(gs)
    big = mb.templates.grid_cluster(6, 3)
    big[1] = mb.Ment("Z")
    with pytest.raises(NotImplementedError):
        lower(big, mixed=True)
    gs[1] = mb.Ment((0.1, 0.2), "XYZ")  # fixed two-angle XYZ: lowered on the density-matrix path (ment.py:239-251)
    xyz = [st for st in lower(gs, mixed=True).steps if st.node == 1][0]
    n = np.array([np.cos(0.1) * np.cos(0.2), np.sin(0.1) * np.cos(0.2), np.sin(0.2)])
    assert xyz.plane == _lib.PLANE_XYZ and np.allclose([xyz.fixed_cos, xyz.fixed_sin, xyz.fixed_z], n, atol=1e-16)
    with pytest.raises(ValueError, match="only XY plane is supported"):
        lower(gs)  # the state-vector path rejects it like every non-XY plane
    gs[1] = mb.Ment("XYZ")  # trainable XYZ: one float arrives where the reference needs a tuple
    with pytest.raises(TypeError, match="Expected tuple"):
        lower(gs, mixed=True)
    tri = mb.MBQCircuit(mb.GraphState([(0, 1), (1, 2), (2, 0)]), input_nodes=[0], output_nodes=[2])
    with pytest.raises(ValueError, match="Schedule must be provided"):
        lower(tri)
    assert not window_is_valid(lower(mb.templates.muta(2, 1, one_column=True)))
    assert window_is_valid(lower(mb.templates.muta(2, 1, one_column=True), window_size=4))


def test_noise_block_coefficients():
    nz = noise_from_kraus(kraus_set("depolarizing", p=0.3))
    assert np.allclose(list(nz.pop), [1 - 0.2, 0.2, 0.2, 1 - 0.2])
    assert np.isclose(nz.coh_g, 1 - 0.4) and np.isclose(nz.coh_d, 0.0)
    nz = noise_from_kraus(kraus_set("amplitude_damping", p=0.25))
    assert np.allclose(list(nz.pop), [1, 0.25, 0, 0.75]) and np.isclose(nz.coh_g, np.sqrt(0.75))
    nz = noise_from_kraus(kraus_set("bit_flip", p=0.1))
    assert np.allclose(list(nz.pop), [0.9, 0.1, 0.1, 0.9]) and np.isclose(nz.coh_d, 0.1)
    with pytest.raises(ValueError):
        kraus_set("nope")
    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)  # unitary mixing populations into coherences
    with pytest.raises(NotImplementedError):
        noise_from_kraus([h])


def test_facade_contract_without_gpu():
    gs = mb.templates.linear_cluster(3)
    with pytest.raises(ValueError, match="not supported"):
        mb.PatternSimulator(gs, backend="numpy-sv")
    with pytest.raises(TypeError):
        mb.PatternSimulator(gs, input_state=[1, 0])
    with pytest.raises(NotImplementedError):
        mb.PatternSimulator(gs, backend="cuda-sv-stream", force0=False)
    with pytest.raises(NotImplementedError):  # sampled runs: register kernels only
        mb.PatternSimulator(mb.templates.grid_cluster(6, 3), backend="cuda-dm", force0=False)  # DM shots: window <= 5
    assert mb.PatternSimulator(mb.templates.grid_cluster(6, 3), backend="cuda-sv", force0=False).window_size == 7  # SV shots: <= 12
    with pytest.raises(NotImplementedError):
        mb.PatternSimulator(mb.templates.grid_cluster(13, 3), backend="cuda-sv", force0=False)
    assert mb.PatternSimulator(gs, backend="cuda-sv", force0=False, seed=5).seed == 5
    ps = mb.PatternSimulator(gs, backend="CUDA-SV", some_unknown_kwarg=3)
    assert ps.window_size == 2 and ps.mbqcircuit is gs and ps.outcomes == {}
    assert ps.schedule_measure == [0, 1]
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ps.run([0.0, 0.0])


def test_lowering_accepts_reference_circuit_objects():
    """the CUDA backends read only public attributes, so a mentpy.MBQCircuit can be
    handed over unchanged (build container only: needs the reference tree)."""
    from oracle.ref_shim import import_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present")
    mp = import_reference()
    for name, args, w in (("grid_cluster", (2, 6), 1), ("grid_cluster", (3, 5), 6), ("muta", (2, 1), 5)):
        ref = getattr(mp.templates, name)(*args)
        mine = getattr(mb.templates, name)(*args)
        if name == "grid_cluster":
            ref[1] = mp.Ment("X")
            mine[1] = mb.Ment("X")
        a, b = lower(ref, window_size=w), lower(mine, window_size=w)
        assert a.schedule == b.schedule and a.input_slot == b.input_slot and a.output_slot == b.output_slot
        assert a.init_cz_mask == b.init_cz_mask and a.n_angles == b.n_angles
        for sa, sb in zip(a.steps, b.steps):
            assert (sa.node, sa.slot, sa.angle_idx, sa.plane, sa.append, sa.new_node, sa.nbr_mask,
                    sa.fixed_cos, sa.fixed_sin) == (sb.node, sb.slot, sb.angle_idx, sb.plane, sb.append,
                                                    sb.new_node, sb.nbr_mask, sb.fixed_cos, sb.fixed_sin)
        sim = mb.simulators.CudaSimulatorSV(ref, None, window_size=w)   # constructs without a GPU
        assert sim.window_size == a.window


@pytest.mark.parametrize("seed", range(100))
def test_lowered_plan_emulation_matches_oracle_on_random_patterns(seed):
    """The lowered plan (slots, sign masks, input / output slot lists), executed by a numpy
    emulation of the kernels' index arithmetic, against the oracle on the fuzz patterns -- catches
    lowering errors (e.g. the SV output-order rule for merged circuits) without a GPU."""
    from fuzz_patterns import random_pattern
    from oracle import matrix_free
    from plan_emulator import run_dm, run_sv

    mixed = seed % 2 == 1
    gs, w, ang, inp = random_pattern(mb, seed, mixed)
    pat = PatternData.from_circuit(gs)
    pl = lower(gs, window_size=w, mixed=mixed)
    if mixed:
        want, woc = matrix_free.run_dm_batch(pat, ang[None], input_states=inp[None], window_size=w, return_outcomes=True)
        got, oc = run_dm(pl, ang, inp)
        assert np.abs(got - want[0]).max() < 1e-10 and oc == [int(o) for o in woc[0]]
    else:
        want = matrix_free.run_sv_batch(pat, ang[None], input_states=inp[None], window_size=w)[0]
        assert 1 - abs(np.vdot(run_sv(pl, ang, inp), want)) ** 2 < 1e-10


def test_entry_points_validate_arguments_before_touching_cuda():
    """No GPU needed: every batch entry point rejects a NULL plan / bad sizes with MBQC_E_ARG and a
    message, without making a CUDA call."""
    lib = _lib.load()
    null = None
    calls = {
        "mbqc_run_batch_sv": (null, null, 0, null, 0, 4, null, 0, null, null),
        "mbqc_run_batch_sv_f32": (null, null, 0, null, 0, 4, null, 0, null, null),
        "mbqc_run_batch_dm": (null, null, 0, null, 0, 4, null, null, null, null),
        "mbqc_run_batch_dm_expect": (null, null, 0, null, 0, 4, null, null, null, null, null),
        "mbqc_psr_grad_batch": (null, null, 0, null, 0, 4, null, 1.5, null, null, null, null),
        "mbqc_psr_grad_dataset": (null, null, 0, null, null, 1, 1, 1.5, null, null, null, null, null),
        "mbqc_run_batch_sv_sampled": (null, null, 0, null, 0, 4, 1, 0, 0, 1, null, null, null, null, null, null),
        "mbqc_run_batch_dm_sampled": (null, null, 0, null, 0, 4, 1, 0, 0, 1, null, null, null, null, null, null),
        "mbqc_train_dataset": (null, null, null, null, 1, 1, 1.5, null, 0, 1, null, null, null, null, null),
        "mbqc_plan_set_feedforward": (null, null, 0),
    }
    for name, args in calls.items():
        rc = getattr(lib, name)(*args)
        assert rc == _lib.MBQC_E_ARG, name
        assert b"NULL" in lib.mbqc_last_error() or b"plan" in lib.mbqc_last_error(), name
    ticket = C.c_int32(-1)
    rc = lib.mbqc_run_batch_sv_host_submit(null, null, 0, null, 0, 4, null, 0, null, 0, 0, C.byref(ticket))
    assert rc == _lib.MBQC_E_ARG
    assert lib.mbqc_host_wait(10**6, None) == _lib.MBQC_E_ARG
    assert lib.mbqc_psr_grad_dataset_workspace_bytes(None, 1, 1) == -1
    assert lib.mbqc_host_workspace_bytes(None, 1, 0) == -1


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under mentpy_b200/ may import it or /root/reference, and
    in bench.py only the cpu_baseline / --impl reference legs may."""
    pkg = os.path.join(ROOT, "mentpy_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference|ref_shim", re.M)
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not pat.search(text), os.path.join(dirpath, f)
    bench = open(os.path.join(ROOT, "bench.py"), encoding="utf-8").read()
    uses = [m.start() for m in re.finditer(r"^\s*(from|import)\s+oracle\b", bench, re.M)]
    assert uses, "bench.py lost its cpu_baseline leg"
    for pos in uses:  # every import sits inside the CPU-port helpers
        head = bench[:pos]
        fn = re.findall(r"^def (\w+)\(", head, re.M)[-1]
        assert fn in ("_cpu_worker", "_pattern_json"), fn  # the CPU-port worker and its pattern description
