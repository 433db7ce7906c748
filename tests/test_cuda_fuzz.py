"""GPU: the CUDA path against the oracle on the seeded random patterns of fuzz_patterns.py (the
oracle is pinned against the live reference on the very same seeds by test_fuzz_reference_cpu.py)."""
import numpy as np
import pytest

import mentpy_b200 as mb
from conftest import dm_distance
from fuzz_patterns import random_pattern
from oracle import matrix_free
from oracle.pattern_data import PatternData

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(40))
def test_cuda_equals_oracle_on_random_patterns(seed):
    mixed = seed % 2 == 1
    gs, w, ang, inp = random_pattern(mb, seed, mixed)
    pat = PatternData.from_circuit(gs)
    rng = np.random.default_rng(1000 + seed)
    A = np.vstack([ang[None], rng.uniform(0, 2 * np.pi, (20, len(ang)))])
    if mixed:
        ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-dm", window_size=w)
        got, oc = ps.run_batch(A, return_outcomes=True)
        want, woc = matrix_free.run_dm_batch(pat, A, input_states=inp[None], window_size=ps.window_size, return_outcomes=True)
        assert dm_distance(got, want) < 1e-10 and np.array_equal(oc, woc)
        rho = ps.run(ang)
        assert dm_distance(rho, want[0]) < 1e-10
    else:
        ps = mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv", window_size=w)
        got = ps.run_batch(A)
        want = matrix_free.run_sv_batch(pat, A, input_states=inp[None], window_size=ps.window_size)
        assert np.max(1 - np.abs(np.sum(got.conj() * want, axis=1)) ** 2) < 1e-10
        assert np.allclose(got, want, atol=1e-9)           # including the reference's global phase
        # the streaming engine runs the same pattern (single angle set, state in HBM)
        st = mb.PatternSimulator(gs, input_state=inp, backend="cuda-sv-stream", window_size=w)
        psi = st.run(ang, output_form="sv")
        assert 1 - abs(np.vdot(psi, want[0])) ** 2 < 1e-10
