"""Multi-rank K1 steps against the single-GPU call (scripts/check_sharded.py).

The driver's box has ONE GPU: the 2- and 4-rank tests therefore run every rank as its own process on cuda:0 (gloo
rendez-vous, CUDA IPC between processes of the same device, the GPU time-slices the ranks).  That exercises the real
exchange code with world > 1 -- rank offsets, the words pushed into other processes' windows, the polling collect -- which a
world = 1 run cannot.  With >= 2 GPUs the same script also runs over NCCL, one rank per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, port, *flags, timeout=900, knobs=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "check_sharded.py"), *flags]
    env = dict(os.environ)
    env.pop("PH_BATCH4", None)
    env.update(knobs or {})
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)
    dump = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(dump):   # keep the ranks' own output next to the other run artefacts
        open(os.path.join(dump, f"check_sharded_{nproc}{'_'.join(flags)}.log"), "w").write(out.stdout + "\n---- stderr ----\n" + out.stderr)
    mine = "\n".join(l for l in (out.stdout + out.stderr).splitlines() if "rank" in l or "Error" in l or "error" in l)
    assert out.returncode == 0, mine[-4000:] + out.stdout[-1500:]
    assert f"sharded path == single GPU path on {nproc} ranks: True" in out.stdout, out.stdout[-3000:]
    return out.stdout


def test_two_ranks_on_one_gpu():
    out = _run(2, 29533, "--one-gpu")
    assert "snv-sharded peer" in out and "peer" in out


def test_four_ranks_on_one_gpu():
    _run(4, 29535, "--one-gpu", "--quick")


@pytest.mark.parametrize("item_batches", ["2", "3"])
def test_two_ranks_multi_batch_items(item_batches):
    """The pushed LL slots of items of 2 / 3 consecutive 8-line batches (k_pass NB8; chosen automatically only on inputs
    with several batches per warp of the grid, forced here)."""
    _run(2, 29541 + int(item_batches), "--one-gpu", "--quick", knobs={"PH_ITEM_BATCHES": item_batches})


def test_sharded_nccl_equals_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (the one-GPU tests above cover the peer exchange with world > 1)")
    out = _run(2, 29537)
    assert "nccl" in out


def test_snv_sharded_one_rank():
    """The LL push + polling collect of the SNV-sharded step with world = 1 (the window is local); must equal the plain
    single-GPU call bit for bit, on the batched-4 stream (default) and on the flat stream with 8 and 4 lines per batch."""
    import numpy as np
    import torch.distributed as dist

    from phertilizer_b200.sharded import SnvShardedStore
    from phertilizer_b200.store import ObservationStore
    from phertilizer_b200.synth import make_config

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("gloo", rank=0, world_size=1)
    saved = os.environ.get("PH_BATCH4")
    try:
        for name, over, b4 in (("T1", {}, "1"), ("T0", {}, "0"), ("C2", {"n": 3000, "m": 4099}, "1"), ("C2", {"n": 3000, "m": 4099}, "0")):
            os.environ["PH_BATCH4"] = b4
            s = make_config(name, with_embedding=False, **over)
            cell, snv, var, total = (x.numpy() for x in (s.cell, s.snv, s.var, s.total))
            full = ObservationStore.from_coo(s.n, s.m, cell, snv, var, total, device=0)
            sh = SnvShardedStore.from_coo(s.n, s.m, cell, snv, var, total, 0, 1, 0)
            rng = np.random.default_rng(5)
            for trial in range(3):
                group = (8, 4, -1)[trial]          # lines per batch of the flat stream
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
        if saved is None:
            os.environ.pop("PH_BATCH4", None)
        else:
            os.environ["PH_BATCH4"] = saved
        dist.destroy_process_group()
