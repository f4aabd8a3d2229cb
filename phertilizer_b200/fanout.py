"""Independent `--runs` spread across the GPUs of the box (reference loop: phertilizer/run.py:55-74).

Run i of the reference is `ph.phertilize(..., seed=100*seed+i)` on the same `Phertilizer`; `phertilize` re-creates its
RNG, parameters and memo tables on every call (phertilizer.py:408-431), so the runs are independent and can execute
concurrently -- one worker process per device, each with its own device-resident copy of the data.  The winner is chosen
exactly like the sequential loop does: highest normalised log-likelihood, ties to the lowest run index (strict `>`).
"""
from __future__ import annotations

import io
import multiprocessing as mp
import os
from concurrent.futures import ProcessPoolExecutor
from contextlib import redirect_stdout

_PH = None


def device_count() -> int:
    import ctypes as C

    from . import _lib

    n = C.c_int(0)
    _lib.check(_lib.lib().ph_device_count(C.byref(n)))
    return int(n.value)


def visible_devices():
    """PHERTILIZER_DEVICES="0,1,3" selects devices; default: every CUDA device of the box."""
    env = os.environ.get("PHERTILIZER_DEVICES")
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    return list(range(device_count()))


_NAMES = ("cell_labels", "mut_labels", "cell_idx", "mut_idx", "var", "total", "bins")


def _init(queue, shared_dir, ctor_kwargs):
    global _PH
    import numpy as np

    from .phertilizer import Phertilizer

    device = queue.get()
    # the parent indexed the observations once (phertilizer.prepare_inputs) and left the arrays in shared memory: the workers
    # map them instead of unpickling two DataFrames and repeating the label pass each
    arrays = []
    for name in _NAMES:
        path = os.path.join(shared_dir, name + ".npy")
        labels = name.endswith("_labels")
        arrays.append(np.load(path, allow_pickle=labels, mmap_mode=None if labels else "r"))
    _PH = Phertilizer(None, None, device=device, _indexed=tuple(arrays), **ctor_kwargs)


def _one(args):
    i, kwargs, quiet = args
    buf = io.StringIO()
    if quiet:
        with redirect_stdout(buf):
            out = _PH.phertilize(**kwargs)
    else:
        out = _PH.phertilize(**kwargs)
    return i, out, buf.getvalue(), (_PH.get_id_mappings() if i == 0 else None)


def run_many(variant_data, bin_count_data, ctor_kwargs, run_kwargs, devices=None, quiet=True):
    """run_kwargs: list of `phertilize` keyword dicts (one per run).  Returns ([(tree, tree_list, loglikes, stdout)] in
    run order, (cell_lookup, mut_lookup)).  One worker per entry of `devices` (a device may be listed more than once)."""
    devices = visible_devices() if devices is None else list(devices)
    if not devices:
        raise RuntimeError("no CUDA device: phertilizer_b200 has no CPU fallback")
    import shutil
    import tempfile

    import numpy as np

    from .phertilizer import prepare_inputs

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    for d in devices:
        queue.put(d)
    results = [None] * len(run_kwargs)
    lookups = None
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    shared_dir = tempfile.mkdtemp(prefix="phertilizer_b200_", dir=base)
    try:
        for name, arr in zip(_NAMES, prepare_inputs(variant_data, bin_count_data)):
            np.save(os.path.join(shared_dir, name + ".npy"), np.asarray(arr), allow_pickle=name.endswith("_labels"))
        with ProcessPoolExecutor(max_workers=min(len(devices), len(run_kwargs)), mp_context=ctx, initializer=_init,
                                 initargs=(queue, shared_dir, ctor_kwargs)) as pool:
            for i, out, text, lk in pool.map(_one, [(i, kw, quiet) for i, kw in enumerate(run_kwargs)]):
                results[i] = (*out, text)
                if lk is not None:
                    lookups = lk
    finally:
        shutil.rmtree(shared_dir, ignore_errors=True)
    return results, lookups


def best_run(results):
    """run.py:71-73: strict `>` on the tree's normalised log-likelihood, first (lowest index) wins ties."""
    best, best_like = None, float("-inf")
    for i, r in enumerate(results):
        if r[0].norm_loglikelihood > best_like:
            best_like, best = r[0].norm_loglikelihood, i
    return best
