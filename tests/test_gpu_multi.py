"""GPU, >= 2 devices: the iterated mode and the row-sharded SpMV across REAL GPUs -- one process per
device, CUDA IPC peer mappings, the library's own NCCL communicator -- against the CPU oracle
(tests/multi_worker.py is one rank).  Skipped on boxes with a single GPU; the same logic runs there
with emulated ranks (tests/test_gpu_synth.py) and on CPU over gloo (tests/test_distributed_cpu.py)."""
import ctypes
import json
import socket
import subprocess
import sys
from pathlib import Path

import pytest

from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def n_devices() -> int:
    n = ctypes.c_int(0)
    try:
        return n.value if load_package().lib().b200_get_device_count(ctypes.byref(n)) == 0 else 0
    except Exception:
        return 0


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_world(world: int, tmp_path):
    port = free_port()
    outs = [tmp_path / f"rank{r}.json" for r in range(world)]
    procs = [subprocess.Popen([sys.executable, str(ROOT / "tests" / "multi_worker.py"), str(r), str(world), str(port),
                               str(outs[r])], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            pytest.fail("multi-GPU worker timed out")
    for r, p in enumerate(procs):
        assert outs[r].exists(), f"rank {r} died (rc {p.returncode}):\n{logs[r][-4000:]}"
    results = [json.loads(o.read_text()) for o in outs]
    for res in results:
        bad = {k: v for k, v in res["checks"].items() if not v["ok"]}
        assert not bad, f"rank {res['rank']}: {bad}"
        assert len(res["checks"]) >= 30
    for r, p in enumerate(procs):
        assert p.returncode == 0, logs[r][-4000:]
    return results


@pytest.mark.skipif(n_devices() < 2, reason="needs at least 2 GPUs")
def test_iterated_mode_and_sharded_spmv_on_two_real_gpus(tmp_path):
    res = run_world(2, tmp_path)
    assert res[0]["nccl_version"] >= 22000
    print("multicast path:", res[0].get("mcast"))


@pytest.mark.skipif(n_devices() < 4, reason="needs at least 4 GPUs")
def test_iterated_mode_on_four_real_gpus(tmp_path):
    run_world(4, tmp_path)
