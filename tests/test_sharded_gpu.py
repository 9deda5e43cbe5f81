"""GPU test of the multi-process sharding path (reference slaf/distributed/coordinator.py:40-80): N processes
open ONE expression table, each reads its contiguous fragment range through shard_source + SLAFIterableDataset on a GPU, and the
union of the emitted cells is every cell exactly once with bit-exact tokens.  Uses every visible GPU round robin (two ranks
share the device on a one-GPU box: the sharding logic and the per-rank contexts are what is under test, not the fabric)."""

import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("route", ["csi", "coo"])
@pytest.mark.parametrize("world", [2, 3])
def test_fragment_range_shards_emit_every_cell_once(tmp_path, world, route):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_sharded_worker.py"), str(tmp_path), route]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(tmp_path / "result.json"))
    assert len(res) == world
    all_ids = [c for rk in res for c in rk["ids"]]
    assert sorted(all_ids) == list(range(1500)), "every cell exactly once over the ranks"
    assert all(rk["tokens_ok"] for rk in res)
    assert all(rk["used_csi"] == (route == "csi") for rk in res)
    bounds = sorted(tuple(rk["rows"]) for rk in res)
    assert bounds[0][0] == 0 and all(a[1] == b[0] for a, b in zip(bounds[:-1], bounds[1:])), "row ranges tile the table"
    assert all(len(rk["ids"]) > 0 for rk in res)
