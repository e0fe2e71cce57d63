"""z-slab decomposed solver on >= 2 GPUs against the single-GPU solver (skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_decomposition_matches_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29517', os.path.join(ROOT, 'tests', 'mgpu_parity.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420)       # the script bounds itself at 240 s (watchdog)
    assert 'MGPU_PARITY_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
