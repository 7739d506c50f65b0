"""The opt-in send/recv + allreduce exchange transport (TOE_DIST_XCHG=sendrecv — the default of round 1, see csrc/dist.cu:xchg_mode for
why it no longer is) on real GPUs: same parity bars as the default all-gather transport (tests/test_gpu_y_dist.py), through
tests/dist_worker.py under torchrun; skipped when fewer than 2 GPUs are visible."""
import pytest

from test_dist import _ngpus, _torchrun

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims,extra", [("24,8,4", []), ("48,16,6", ["simp"])])
def test_two_gpu_sendrecv_transport(dims, extra):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, ["tests/dist_worker.py", dims] + extra, 29541, env={"TOE_DIST_XCHG": "sendrecv", "TOE_EXPECT_TRANSPORT": "nccl"})
    assert r.returncode == 0 and "DIST PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_four_gpu_sendrecv_transport():
    if _ngpus() < 4:
        pytest.skip("needs 4 GPUs")
    r = _torchrun(4, ["tests/dist_worker.py", "48,16,6"], 29542, timeout=400, env={"TOE_DIST_XCHG": "sendrecv", "TOE_EXPECT_TRANSPORT": "nccl"})
    assert r.returncode == 0 and "DIST PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
