"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a 1-GPU box): spawns torchrun on tests/dp_check.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("peer", [False, True])
def test_batch_sharded_matches_single_gpu(peer):
    """peer=False: mu via NCCL all-reduce; peer=True: all-reduce fused into the adjoint kernel over NVLink peer memory."""
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29611" if not peer else "29613", os.path.join(HERE, "dp_check.py")] + \
          (["--peer"] if peer else [])
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "dp_check ok" in p.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_state_on_another_gpu_than_the_current_one_is_refused():
    """The kernels launch on the current device's stream: a state tensor of another GPU must fail loudly, not race."""
    from pnode import petsc_adjoint
    from pnode_b200.errors import Error
    from _problems import SpiralFunc, spiral_inputs

    torch.cuda.set_device(0)
    u0 = spiral_inputs(8)[0].to("cuda:1")
    ode = petsc_adjoint.ODEPetsc()
    with pytest.raises(Error, match="current CUDA device"):
        ode.setupTS(u0, SpiralFunc().to("cuda:1"), step_size=0.025, method="rk4", enable_adjoint=True)
