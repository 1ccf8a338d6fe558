"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py): seeded inputs of one `derivs` with the outputs of the CPU
oracle frozen at the time they were made.  They are not NDSPMHD outputs (the reference cannot be built here).

* CPU: the oracle of today must reproduce them (a drift alarm for the checker itself), and the generators must still produce the inputs.
* GPU (tests/test_gpu_vectors.py): the CUDA path through the C-ABI is compared with the frozen outputs, without executing anything
  under oracle/.
"""
import glob
import json
import os

import numpy as np
import pytest

import parity
from ndspmhd_b200 import abi, setups  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_case(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    ndim, npart, nt = int(d["ndim"]), int(d["npart"]), int(d["ntotal"])
    o = abi.NdOptions.from_buffer_copy(d["opts"].tobytes())
    idim = nt + 64
    pin, pout = abi.Particles(ndim, npart, idim), abi.Particles(ndim, npart, idim, nt)
    for k in pin.arrays:
        pin.arrays[k][:npart] = d["in_" + k]
        pout.arrays[k][:nt] = d["out_" + k]
    scal = json.loads(d["scalars"].tobytes().decode())
    return o, pin, pout, scal, int(d["aux"])


def test_fixtures_exist():
    assert {"briowu1d", "ot2d_closepacked", "ot3d_glass_fast", "dustybox3d"} <= set(NAMES)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_its_golden_outputs(name):
    from oracle import oracle
    o, pin, pout, scal, aux = load_case(name)
    assert bytes(o) == bytes(abi.NdOptions.from_buffer_copy(bytes(o)))
    p = pin.copy()
    s, _ = oracle.derivs(o, p)
    errs = parity.assert_parity(p, pout, s, scal, o, aux=bool(aux), rtol=1e-13)
    assert max(errs.values()) <= 1e-13


def test_generators_still_make_the_golden_inputs():
    """The seeded setups feed every other parity test: their inputs must not drift either."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for name in NAMES:
        if name not in mg.CASES:
            continue
        o, pin, pout, scal, aux = load_case(name)
        make, aux2 = mg.CASES[name]
        o2, p2 = make()
        o2.device_ghosts = 1
        o2.want_aux = aux2
        assert aux == aux2 and bytes(o2) == bytes(o), name
        assert p2.npart == pin.npart
        for k in pin.arrays:
            assert np.array_equal(p2.arrays[k][: pin.npart], pin.arrays[k][: pin.npart]), (name, k)
