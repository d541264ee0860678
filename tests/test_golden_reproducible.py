"""The committed fixtures are exactly what tests/golden/make_golden.py produces from the unmodified reference
(/root/reference, compiled into oracle/_ref) - bit for bit. Skipped where the reference tree is absent (the GPU box)."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not os.path.isdir("/root/reference/solver"), reason="needs the reference tree")
def test_fixtures_regenerate_bit_for_bit():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    mk.build()
    for fresh, path in ((mk.parse_scripts(), "reference_scripts.npz"), (mk.reference_outputs(), "reference_outputs.npz"),
                        (mk.reference_outputs_msg(), "reference_outputs_msg.npz")):
        committed = np.load(os.path.join(HERE, "golden", path))
        assert set(fresh.keys()) == set(committed.files)
        for key in committed.files:
            a, b = np.asarray(fresh[key]), committed[key]
            assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), key
