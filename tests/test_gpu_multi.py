"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tests/mgpu/sweep_check.py with one process
per GPU through torch.distributed.run -- every rank checks the whole gathered result of the fused P2P gather
(frx_history_gather behind sweep.history_sweep) and of the host-result sweeps against the CPU oracle."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('world', [2])
def test_sweeps_on_several_gpus_every_rank_checks_everything(world):
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(REPO, 'tests', 'mgpu', 'sweep_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO)
    ok = [ln for ln in r.stdout.splitlines() if ln.startswith('SWEEP_CHECK_OK')]
    assert r.returncode == 0 and len(ok) == world, (r.stdout[-2000:], r.stderr[-4000:])
    assert all('fused_p2p_gather=True' in ln for ln in ok), ok
