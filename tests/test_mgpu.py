"""Multi-GPU bit-exactness, collected by pytest (VERDICT r1 'what's weak' 2): spawns tests/mgpu_check.py under torchrun
with one rank per GPU when the box has at least two -- every rank's merged filters, count tables, flags and qualities
against the oracle on the whole input, with both merge transports."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _gpus() -> int:
    from kbbq_b200 import api

    return api.lib().kbbq_device_count()


@pytest.mark.parametrize("merge", ["nvls", "nvls-unavailable", "p2p", "allgather"])
def test_sharded_run_matches_the_oracle_on_every_rank(merge):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least two GPUs on the box")
    world = 2 if n < 4 else 4
    env = dict(os.environ, KBBQ_MERGE=merge.split("-")[0], MASTER_ADDR="127.0.0.1")
    if merge.startswith("nvls"):  # the switch ORs the copies (filters as shareable VMM memory); where that is not to be had --
        env["KBBQ_NVLS_VERBOSE"] = "1"  # simulated by "nvls-unavailable" -- the filters are re-created and the peer-memory merge runs
    if merge == "nvls-unavailable":
        env["KBBQ_NVLS_FAIL"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", {"p2p": "29517", "allgather": "29518", "nvls": "29519", "nvls-unavailable": "29520"}[merge], os.path.join(ROOT, "tests", "mgpu_check.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    log = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log):
        with open(os.path.join(log, f"mgpu_check_{merge}_w{world}.log"), "w") as fh:
            fh.write(p.stdout + "\n--- stderr ---\n" + p.stderr)
    assert p.returncode == 0 and "MGPU PARITY OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    if merge == "nvls-unavailable":
        assert "merge=p2p" in p.stdout or "merge=allgather" in p.stdout
