"""ROWS assembly variant on a B200 (file name sorts last on purpose: the variant is not the default yet)."""
import pytest

import rows_variant_checks as rc

pytestmark = pytest.mark.gpu


def test_rows_variant(pkg, fo, golden_c1):
    ctx = pkg.Context(0)
    try:
        rc.check_rows_variant(pkg, fo, ctx, golden_c1)
    finally:
        ctx.close()
