"""The SpMV's bulk-copy / mbarrier pipeline under LATE, OUT-OF-ORDER completion of the copies (EMU_BULK_DELAY, tests/cuda_emu): on
hardware the fills of consecutive stages may land in any order; a protocol whose parity waits can be satisfied by the phase before
the one they mean then reads a stage before its data has arrived.  That was the root cause of the transient CG faults of round 1 (about
2 in 10^6 launches at 10M tets): the two consumer groups took alternate chunks, so the fill of a stage before a group's own belonged to
the other group, and nothing made the group see it complete.  TEST INFRASTRUCTURE (emulated build), not a product path."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu_support  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    pkg, lib = emu_support.load_emu()
    with emu_support.emulated(pkg, lib):
        yield pkg, lib


def _problem(pkg, ctx):
    pts, cells = pkg.meshgen.cantilever(24, 8, 4)                      # 1125 nodes = 18 chunks of 64 rows
    ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
    lam, mu = pkg.create_material_model(1.0, 0.3)
    ctx.assemble_lame(lam, mu)
    load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0); fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
    ctx.add_nodal_force(load, [0.0, 0.0, -1.0])
    nfd = ctx.node_dofs()
    pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
    ctx.apply_dirichlet(pres)
    return pres


@pytest.mark.parametrize("grid", ["1", "2", "5"])
def test_spmv_pipeline_with_out_of_order_fills(emu, monkeypatch, grid):
    pkg, lib = emu
    ctx = pkg.Context(0)
    try:
        pres = _problem(pkg, ctx)
        x = np.random.default_rng(11).standard_normal(ctx.ndofs); x[pres - 1] = 0.0
        y_ref = ctx.spmv(x)
        monkeypatch.setenv("TOE_SPMV_GRID", grid)                      # few CTAs: every CTA goes through several laps of its 3 stages
        assert np.array_equal(ctx.spmv(x), y_ref)
        for delay in ("3", "17", "60"):                                # scheduler passes a late fill is held back
            monkeypatch.setenv("EMU_BULK_DELAY", delay)
            for _ in range(3):                                         # the per-fill delays differ from launch to launch
                assert np.array_equal(ctx.spmv(x), y_ref), (grid, delay)
    finally:
        ctx.close()


def test_pcg_with_out_of_order_fills_is_bit_identical(emu, monkeypatch):
    pkg, lib = emu
    ctx = pkg.Context(0)
    try:
        _problem(pkg, ctx)
        monkeypatch.setenv("TOE_SPMV_GRID", "2")                       # same grid in both solves: the fused p'Ap sums one partial per CTA
        st0 = ctx.solve_pcg(1e-8, 1e-8, 2000, graph=False, history=True)
        u0 = ctx.solution()
        monkeypatch.setenv("EMU_BULK_DELAY", "25")
        st1 = ctx.solve_pcg(1e-8, 1e-8, 2000, graph=False, history=True)
        assert st0["converged"] == 1 and st1["converged"] == 1 and st1["niter"] == st0["niter"]
        assert np.array_equal(st1["residuals"], st0["residuals"]) and np.array_equal(ctx.solution(), u0)
    finally:
        ctx.close()


def test_partitioned_pcg_with_out_of_order_fills(emu, monkeypatch):
    """the single-reduction CG of the partitioned path — the one that could not recover from a stale chunk — on 2 emulated ranks"""
    import test_emu_dist as ed
    pkg, lib = emu
    prob = ed._problem(pkg, (16, 6, 3), False)
    monkeypatch.setenv("TOE_SPMV_GRID", "2")
    ref = ed._run_ranks(pkg, 2, prob, False, repeats=1, tol=1e-9)
    monkeypatch.setenv("EMU_BULK_DELAY", "25")
    got = ed._run_ranks(pkg, 2, prob, False, repeats=1, tol=1e-9)
    for rk in range(2):
        a, b = ref[rk][0], got[rk][0]
        assert b["conv"] == 1 and b["brk"] == 0 and a["it"] == b["it"] and np.array_equal(a["u"], b["u"]), (rk, a["it"], b["it"])
