"""N > 1 on real GPUs (skipped on a single-GPU box): the sharded loop query with the fused one-kernel
exchange, the push exchange and NCCL must all equal the CPU oracle."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_loop_query_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the world_size-2 exchange is covered over gloo in test_host_logic.py)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "multi_gpu_check.py"), "60000"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["push_equals_nccl"] and d["equals_oracle"]
    assert d["fused"]["exchange"] == "fused" and d["fused"]["best_is_source"]
