"""Multi-GPU path on real devices (needs >= 2 GPUs; `gpurun --gpus 2`): one
process per GPU, disjoint MWC seed sets, ONE NCCL all-reduce of the 64-bit
accumulators.  The host logic alone is covered on CPU by test_parallel_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        from pyxopto_b200.cu import abi
        return abi.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason='needs at least 2 GPUs')
def test_nccl_allreduce_equals_sum_of_shards():
    world = min(_gpu_count(), 2)
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests', 'multi_gpu_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert 'MULTI_GPU_RESULT OK' in out.stdout
    assert 'MULTI_GPU_SWEEP OK' in out.stdout
