"""N>1 path, CPU part: the host-side plumbing (unique-id broadcast, rank-0-only reference arm) with world_size-2 gloo process
groups.  The GPU part (partition invariance through tests/dist_worker.py under torchrun) lives in tests/test_gpu_y_dist.py, which
sorts after the single-GPU parity files; `_torchrun` / `_ngpus` below are shared with it."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script_args, port, timeout=280, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    e = dict(os.environ); e.update(env or {})
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=e)


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


# ---------------------------------------------------------------------------------------------------------------
# CPU (gloo, world_size 2)
# ---------------------------------------------------------------------------------------------------------------
GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import __graft_entry__ as graft
pkg = graft.load_package()
dist.init_process_group("gloo")
rank, local_rank, world = pkg.parallel.env_rank()
assert world == 2 and dist.get_rank() == rank
payload = bytes(range(128)) if rank == 0 else None
got = pkg.parallel.broadcast_bytes(dist, payload, 128, 0)
assert got == bytes(range(128)), got
# every rank derives the same clamp/load sets and the same mesh from the same generator (no data exchange needed)
pts, cells = pkg.meshgen.cantilever(6, 3, 2)
t = torch.tensor([float(pts.sum()), float(cells.sum())], dtype=torch.float64)
lst = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(lst, t)
assert torch.equal(lst[0], lst[1])
# without a GPU the product path must refuse loudly on every rank
try:
    pkg.Context(local_rank)
    ok = torch.cuda.is_available()
except pkg.TopOptError:
    ok = True
assert ok
dist.barrier()
print("GLOO OK rank", rank, flush=True)
dist.destroy_process_group()
'''


def test_gloo_world2_host_plumbing(tmp_path):
    script = tmp_path / "gloo_worker.py"
    script.write_text(GLOO_WORKER % {"root": ROOT})
    r = _torchrun(2, [str(script)], 29533, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("GLOO OK") == 2


def test_reference_arm_prints_one_line_under_torchrun():
    """`bench.py --impl reference` launched like the B200 arm (torchrun, N=2): rank 0 alone works and prints."""
    r = _torchrun(2, ["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], 29534, timeout=280,
                  env={"TOE_BENCH_CPU_MESH": "tiny", "TOE_BENCH_CPU_SLICE": "2000", "TOE_BENCH_CPU_ITERS": "5"})
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "elements/s" and d["value"] > 0 and d["n_gpus"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # same metric / config as the B200 arm prints for the same arguments, and the extrapolation is spelled out
    import types
    sys.path.insert(0, ROOT)
    import bench
    assert d["metric"] == bench.METRIC and d["config"] == bench.static_config(types.SimpleNamespace(workload="C4_10M", matrix_free=False, gpus=2))
    ex = d["cpu_baseline"]["extrapolation"]
    assert ex["pcg_iterations"] == 13689 and abs(ex["value"] - d["value"]) < 1e-9 and "Extrapolated" in d["cpu_baseline"]["sample"]
