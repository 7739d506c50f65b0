"""bench.py's B200 arm driven end to end on the EMULATED build (tests/cuda_emu, test infrastructure): the toy workload,
torch.cuda calls stubbed.  Checks the script's control flow and the JSON contract — never a measurement."""
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu_support  # noqa: E402


def _args(**kw):
    d = dict(gpus=1, steps=2, warmup=1, impl="b200", workload="toy", matrix_free=False, no_cpu_baseline=True, no_two_level=False, deadline=600.0, e2e_steps=None, no_e2e_warmup=False, no_l2=False)
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.mark.parametrize("mf", [False, True])
def test_bench_b200_arm_contract_on_emulated_build(monkeypatch, mf):
    import torch
    pkg, lib = emu_support.load_emu()
    import bench
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    args = _args(matrix_free=mf, no_two_level=mf)
    def probe_in_process(a, energy_jacobi, jacobi_s):             # the child process of the real bench, in-process on the emulated build
        b2 = io.StringIO()
        with redirect_stdout(b2):
            bench.two_level_probe(a, pkg)
        d2 = json.loads(b2.getvalue().splitlines()[-1])
        d2["energy_rel_diff_vs_jacobi"] = abs(d2["energy"] - energy_jacobi) / abs(energy_jacobi)
        return d2
    monkeypatch.setattr(bench, "two_level_probe_in_child", probe_in_process)
    buf = io.StringIO()
    with emu_support.emulated(pkg, lib), redirect_stdout(buf):
        bench.run_b200(args, pkg, bench.Progress(args, 0))
    lines = [ln for ln in buf.getvalue().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "clocks", "e2e", "gpu_launches", "roofline", "stages", "metric_parts"):
        assert key in d, key
    assert "error" not in d
    assert d["metric"] == json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "BASELINE.json")))["metric"]
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["unit"] == "elements/s" and d["dtype"] == "f64"
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    assert d["config"] == bench.static_config(args)               # the reference arm prints the very same dict
    e2e = d["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    mp = d["metric_parts"]
    assert mp["elements_assembled_per_s"] is None or mp["elements_assembled_per_s"] > 0
    assert mp["pcg_seconds_to_1e-8"] > 0 and mp["pcg_iterations"] > 0
    assert d["stages"]["pcg_converged"] and d["stages"]["pcg_iterations_per_step"] == [mp["pcg_iterations"]] * 2 and d["stages"]["energy"] > 0
    l2 = d["stages"]["l2_criterion"]
    assert l2["converged"] and l2["pcg_iterations"] != mp["pcg_iterations"] and l2["rel_res_l2"] < 1e-7
    tl = d["stages"]["two_level_preconditioner"]
    assert (mf and tl is None) or "error" not in tl and tl["converged"] and tl["pcg_iterations"] < mp["pcg_iterations"] and tl["energy_rel_diff_vs_jacobi"] < 1e-6


def test_bench_fails_fast_with_one_error_line(monkeypatch):
    """A step that does not converge ends the run at once: one JSON line with "error" and the stage, non-zero exit — no retries."""
    import torch
    pkg, lib = emu_support.load_emu()
    import bench
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.setattr(bench, "ITMAX", 3)                        # cannot converge
    codes = []

    class Exit(Exception):
        pass

    def fake_exit(code):
        codes.append(code)
        raise Exit()
    monkeypatch.setattr(bench.os, "_exit", fake_exit)
    args = _args(no_two_level=True)
    buf = io.StringIO()
    with emu_support.emulated(pkg, lib), redirect_stdout(buf), pytest.raises(Exit):
        bench.run_b200(args, pkg, bench.Progress(args, 0))
    lines = [ln for ln in buf.getvalue().splitlines() if ln.startswith("{")]
    assert codes == [4] and len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] is None and "did not converge" in d["error"] and d["stage"].startswith("warm-up 0") and d["partial"]["niter"] == 3


def test_bench_partitioned_flow_on_emulated_build(monkeypatch):
    """The N > 1 control flow of bench.py cannot meet real NCCL here; it is driven with WORLD_SIZE=2 against ONE emulated rank (a 1-rank
    communicator, torch.distributed stubbed) so that every statement of that path executes at least once."""
    import torch
    import torch.distributed as tdist
    pkg, lib = emu_support.load_emu()
    import bench
    monkeypatch.setenv("WORLD_SIZE", "2"); monkeypatch.setenv("RANK", "0"); monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{kk: vv for kk, vv in k.items() if kk != "device"}))
    for name in ("init_process_group", "barrier", "all_reduce", "destroy_process_group"):
        monkeypatch.setattr(tdist, name, lambda *a, **k: None)

    def one_rank_context(dist, local_rank):
        ctx = pkg.Context(local_rank)
        ctx.comm_init(1, 0, pkg.Context.comm_unique_id())
        return ctx
    monkeypatch.setattr(pkg.parallel, "create_distributed_context", one_rank_context)
    args = _args(gpus=2, steps=1, no_two_level=True)
    buf = io.StringIO()
    with emu_support.emulated(pkg, lib), redirect_stdout(buf):
        bench.run_b200(args, pkg, bench.Progress(args, 0))
    lines = [ln for ln in buf.getvalue().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert "error" not in d and d["n_gpus"] == 2 and d["config"]["parallelism"] == "dd2"
    assert "cpu_baseline" not in d and d["stages"]["local_sizes"] is not None and d["e2e"]["value"] > 0
