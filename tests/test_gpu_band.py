"""Row-band multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2`): the banded solve
(halo exchange + scalar all-reduce over NCCL) must equal the single-GPU solve to
reduction-order noise and be bit-reproducible for a fixed world size."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_band(world, nx, ny, max_disp=16, extra=()):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "band_worker.py"),
           str(nx), str(ny), str(max_disp), *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("BAND_RESULT ")]
    assert line, out.stdout[-2000:] + out.stderr[-4000:]
    return json.loads(line[0][len("BAND_RESULT "):])


@pytest.mark.parametrize("world,shape", [(2, (1500, 1100)), (2, (4000, 3000)), (4, (4000, 3000)), (8, (6000, 6000))])
def test_banded_equals_single_gpu(world, shape):
    r = run_band(world, *shape)
    assert r["repro"] and r["its_equal"]
    assert r["du_mean"] < 1e-4 and r["dv_mean"] < 1e-4 and r["du_max"] < 2e-3 and r["dv_max"] < 2e-3
    assert r["nav_max"] <= 2


@pytest.mark.parametrize("world,shape", [(2, (1500, 1100)), (4, (4000, 3000))])
def test_banded_first_guess_equals_single_gpu(world, shape):
    """-firstguess -lambdac on row bands (octane_variational_flow_band_fg_dev): the hint field of every level is built
    locally from the band's overlap rows of the first guess"""
    r = run_band(world, *shape, extra=("fg",))
    assert r["repro"] and r["its_equal"]
    assert r["du_mean"] < 1e-4 and r["dv_mean"] < 1e-4 and r["du_max"] < 2e-3 and r["dv_max"] < 2e-3
