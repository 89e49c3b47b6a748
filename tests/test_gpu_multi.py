"""GPU, >= 2 devices: NCCL + real ranks.  The z-slab path end to end through the C-ABI (cub_count_async,
cub_comm_exchange_counts, cub_emit_async, cub_comm_gather_mesh), one process per GPU, compared bit for bit
with the CPU oracle's mesh of the whole volume.  Skipped on a single-GPU box (NCCL refuses two ranks on one
device); `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py` runs it (log under profiles/)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case", ["smooth64_tri_proj", "smooth_u8_quads_celldata", "gyroid_tri_noproj", "smooth_i16_tri_proj_celldata"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_gathered_mesh_of_n_ranks_equals_the_oracle(world, case, tmp_path):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "workers", "mgpu_worker.py"), str(tmp_path), case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    verdict = open(os.path.join(tmp_path, "verdict")).read()
    assert verdict.startswith("ok "), verdict
    digests = {open(os.path.join(tmp_path, f"digest{r}")).read() for r in range(world)}
    assert len(digests) == 1, "the ranks did not receive the same gathered mesh"
