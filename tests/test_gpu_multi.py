"""Row-sharded chain over 2 GPUs (NCCL all-gather of the per-table statistics each sweep) against the
one-GPU chain.  Needs two GPUs; run with `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi`."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["nccl", "p2p", "p2p-incremental"])
def test_two_gpu_row_sharded_chain_matches_one_gpu(transport):
    """transport: the once-per-sweep packet exchange through ncclAllGather, or stored straight into the peers' memory
    over NVLink (csrc/mv_exchange.cu)."""
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", {"nccl": "29517", "p2p": "29518"}.get(transport, "29519"), str(ROOT / "tests" / "mp_gpu_shard.py")]
    env = dict(os.environ, MVG_TEST_P2P="0" if transport == "nccl" else "1", MVG_TEST_INCR="1" if transport.endswith("incremental") else "0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SHARD_OK" in r.stdout, r.stdout[-2000:]
