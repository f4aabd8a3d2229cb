"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (elkebir-group/phertilizer).

This module is the only place that touches ``/root/reference``.  It exists so that
``tests/golden/make_goldens.py`` (run in the build container, never on the GPU box) can execute the
reference's own Python code to (a) pin the oracle restatement in ``oracle/oracle_np.py`` and
(b) generate the golden vectors committed under ``tests/golden/``.

Nothing in ``phertilizer_b200/`` imports this file, and nothing here is shipped.

The reference needs a small, documented shim set to import on this stack:
  1. numpy>=2 removed ``np.NINF / np.Inf / np.NAN`` (used at import time, phertilizer.py:558,
     linear_split.py:429) -> aliases are re-created before import;
  2. ``import umap`` (phertilizer.py:7) and ``import pygraphviz`` (draw_clonal_tree.py:3) are not
     installed -> stub modules; the reference is only ever run with ``dim_reduce=False``;
  3. Python>=3.11 rejects the mutable dataclass default at clonal_tree.py:1105 -> a patched COPY is
     made under a temp dir (the field is never read); /root/reference itself is read-only.
"""
import importlib
import os
import shutil
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("PHERTILIZER_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "phertilizer"))


_loaded = None


def load_reference():
    """Import the reference package (patched copy) and return the dict of its modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    import numpy as np

    np.NINF = -np.inf
    np.Inf = np.inf
    np.NAN = np.nan

    work = tempfile.mkdtemp(prefix="ph_ref_")
    shutil.copytree(os.path.join(REFERENCE_ROOT, "phertilizer"), os.path.join(work, "phertilizer"))
    ct = os.path.join(work, "phertilizer", "clonal_tree.py")
    src = open(ct).read()
    bad = "ancestral_muts: np.array = np.empty(shape=0, dtype=int)"
    assert bad in src, "reference layout changed; update shim"
    open(ct, "w").write(src.replace(bad, "ancestral_muts: np.array = None"))
    for stub, body in (
        ("umap", "class UMAP:\n    def __init__(self, *a, **k):\n        raise RuntimeError('umap stub')\n"),
        ("pygraphviz", "class AGraph:\n    def __init__(self, *a, **k):\n        raise RuntimeError('pygraphviz stub')\n"),
    ):
        try:
            importlib.import_module(stub)
        except ImportError:
            os.makedirs(os.path.join(work, "stubs", stub), exist_ok=True)
            open(os.path.join(work, "stubs", stub, "__init__.py"), "w").write(body)
    sys.path.insert(0, os.path.join(work, "stubs"))
    sys.path.insert(0, work)
    mods = {}
    for name in ("phertilizer", "linear_split", "branching_split", "clonal_tree", "clonal_tree_list",
                 "utils", "data", "params", "run"):
        mods[name] = importlib.import_module(f"phertilizer.{name}")
    _loaded = mods
    return mods


def example_paths():
    d = os.path.join(REFERENCE_ROOT, "example", "input")
    return os.path.join(d, "variant_counts.tsv"), os.path.join(d, "binned_read_counts.csv")
