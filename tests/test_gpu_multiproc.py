"""GPU: GPV_GATHER with one PROCESS per rank (CUDA IPC mapping of the gathering rank's buffers: gpv_gather_create ->
gpv_gather_attach), the shape bench.py --gpus N runs in.  One GPU per rank when the box has them; with fewer visible devices the
ranks share device 0 (separate processes time-slice the GPU: slower, same code path -- the mailbox polls have a timeout, not a
deadlock).  Gathered streams == the single-call result byte for byte, two epochs, two sessions, no retry."""
import json
import os
import socket
import subprocess
import sys

import pytest

from util import ROOT, mesh_path

pytestmark = pytest.mark.gpu


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("name,l1,l2,world", [("cessna", 64, 4, 2), ("torus", 32, 16, 3), ("cessna", 128, 8, 4), ("block", 48, 2, 8)])
def test_gather_across_processes_equals_single_call(product, tmp_path_factory, tmp_path, name, l1, l2, world):
    ndev = product.lib().gpv_device_count()
    if ndev < world and world > 3:
        pytest.skip("%d ranks on %d device(s): run on a multi-GPU box" % (world, ndev))
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    port = free_port()
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    procs = []
    for r in range(world):
        out = str(tmp_path / ("rank%d.json" % r))
        procs.append((out, subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_gather_worker.py"), str(r), str(world), str(port), str(r % ndev),
                                            path, str(l1), str(l2), out], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    logs = []
    for out, p in procs:
        try:
            log, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for _, q in procs:
                q.kill()
            pytest.fail("multi-process gather hung")
        logs.append(log)
        assert p.returncode == 0, log[-2000:]
    reports = [json.load(open(out)) for out, _ in procs]
    r0 = reports[0]
    assert len(r0["epochs"]) == 4
    for e in r0["epochs"]:
        assert e["ok_l1"] and e["ok_prefix"] and e["ok_l2"] and e["ok_nb"] and e["ok_counts"], e
    # the Level-2 work is shared out: every boundary cell refined by exactly one rank, no rank idle
    for k in range(4):
        shares = [rep["epochs"][k]["n_refined"] for rep in reports]
        assert sum(shares) == r0["epochs"][k]["nb"], shares
        assert min(shares) > 0 or r0["epochs"][k]["nb"] < 50 * world, shares
