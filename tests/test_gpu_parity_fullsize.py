"""Parity at BASELINE.json's full sizes (C3: 1M tets, C4: 10M tets) on a B200, through size-independent properties
(tests/fullsize_properties.py) and the frozen oracle / regression values in tests/golden/fullsize_c3.json."""
import json
import os

import pytest

import fullsize_properties as fp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_c3.json")


@pytest.mark.parametrize("name,simp", [("C3_1M", False), ("C3_1M", True), ("C4_10M", False)])
def test_fullsize_properties(pkg, name, simp):
    dims = pkg.meshgen.SIZES[name]
    golden = None
    if not simp and os.path.exists(GOLDEN):
        golden = json.load(open(GOLDEN)).get(name)
    ctx = pkg.Context(0)
    try:
        out = fp.run_properties(pkg, ctx, dims, simp=simp, golden=golden, check_pattern=(name == "C3_1M" and not simp),
                                itmax=300000 if simp else 40000)
    finally:
        ctx.close()
    assert out["niter"] > 0
