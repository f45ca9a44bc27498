"""Multi-rank NCCL parity ON HARDWARE: `torchrun --nproc-per-node 2 bench.py --gpus 2 --scaling strong` shards the
10 000-user problem over two GPUs and asserts, inside bench.sharded_parity, that the log-evidence and the
global-parameter gradients (all-reduced) equal the unsharded values and that the per-user gradients equal the
corresponding slices.  Skipped on boxes with one GPU (the driver's SCALE runs go through the same assertion)."""
import json
import os
import subprocess
import sys

import pytest
import torch as t

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(t.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("workload", ["cfg2", "cfg5"])
def test_two_rank_strong_scaling_matches_unsharded(workload):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29561", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "3",
           "--workload", workload, "--scaling", "strong", "--no-weak"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    p = line["parity"]
    assert p["lp_rel_err"] < 1e-6 and p["global_grad_rel_err"] < 1e-5 and p["per_user_grad_rel_err"] < 1e-5
