"""Real NCCL path: needs >= 2 GPUs on the box (skipped on the 1-GPU driver run; covered on CPU by test_sharded_gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_nccl_equals_single_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_sharded_nccl.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "== single GPU path on 2 ranks: True" in out.stdout


def test_snv_sharded_records_one_rank():
    """The record push + unpack path of the SNV-sharded step on ONE GPU (world = 1: the window is local), so the driver's
    single-GPU run exercises k_pass<push> and k_peer_collect_rec; must equal the plain single-GPU call bit for bit."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from phertilizer_b200.sharded import SnvShardedStore
    from phertilizer_b200.store import ObservationStore
    from phertilizer_b200.synth import make_config

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        for name, over in (("T1", {}), ("T0", {}), ("C2", {"n": 3000, "m": 4099})):
            s = make_config(name, with_embedding=False, **over)
            cell, snv, var, total = (x.numpy() for x in (s.cell, s.snv, s.var, s.total))
            full = ObservationStore.from_coo(s.n, s.m, cell, snv, var, total, device=0)
            sh = SnvShardedStore.from_coo(s.n, s.m, cell, snv, var, total, 0, 1, 0)
            rng = np.random.default_rng(5)
            for trial in range(3):
                group = (8, 4, -1)[trial]          # lines per batch = lines per pushed record
                full.set_group(group, group)
                sh.store.set_group(group, group)
                lab = np.zeros(s.n + 16, np.uint8)[: s.n]
                lab[:] = rng.integers(0, 3, s.n)
                labA = np.zeros(s.n + 16, np.uint8)[: s.n]
                labA[:] = lab == 1
                l1, n1 = full.assign_linear(labA, float(labA.sum()))
                l2, n2 = sh.assign_linear(labA, float(labA.sum()))
                assert np.array_equal(l1, l2) and np.array_equal(n1, n2), (name, trial)
                b1, m1 = full.assign_branching(lab)
                b2, m2 = sh.assign_branching(lab)
                assert np.array_equal(b1, b2) and np.array_equal(np.asarray(m1).reshape(-1), np.asarray(m2).reshape(-1)), (name, trial)
            assert not sh.peer.timed_out()
            sh.close()
            full.close()
    finally:
        dist.destroy_process_group()
