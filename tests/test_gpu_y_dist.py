"""N>1 path on real GPUs: partition invariance (N-GPU == 1-GPU == oracle) through tests/dist_worker.py under torchrun: the default transport
(one ncclAllGather per exchange) and the peer-memory kernel; skipped when fewer than 2 (4) GPUs are visible.  The file name sorts after
the single-GPU parity files on purpose: under `-x` a surprise here must not hide the single-GPU parity results."""
import pytest

from test_dist import _ngpus, _torchrun


@pytest.mark.gpu
@pytest.mark.parametrize("dims,extra", [("24,8,4", []), ("48,16,6", ["simp"])])
def test_two_gpu_partition_invariance(dims, extra):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, ["tests/dist_worker.py", dims] + extra, 29531, env={"TOE_EXPECT_TRANSPORT": "nccl-allgather"})
    assert r.returncode == 0 and "DIST PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_two_gpu_larger_mesh_is_reproducible():
    """221k tets, three back-to-back partitioned solves per operator: bit-identical u and iteration counts (a race in the
    peer-memory exchange shows up as run-to-run differences), and parity with the single-GPU solve."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, ["tests/dist_worker.py", "96,32,12", "repeat"], 29536, timeout=400, env={"TOE_DIST_P2P": "1", "TOE_EXPECT_TRANSPORT": "peer-memory"})
    assert r.returncode == 0 and "DIST PARITY OK" in r.stdout and "NON-REPRODUCIBLE" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_two_gpu_peer_memory_transport_gives_identical_results():
    """the opt-in fused peer-memory exchange (CUDA IPC mailboxes) must pass the same parity bars as the NCCL transport"""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, ["tests/dist_worker.py", "24,8,4"], 29535, env={"TOE_DIST_P2P": "1", "TOE_EXPECT_TRANSPORT": "peer-memory"})
    assert r.returncode == 0 and "DIST PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_four_gpu_partition_invariance():
    if _ngpus() < 4:
        pytest.skip("needs 4 GPUs")
    r = _torchrun(4, ["tests/dist_worker.py", "48,16,6", "simp"], 29532)
    assert r.returncode == 0 and "DIST PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
