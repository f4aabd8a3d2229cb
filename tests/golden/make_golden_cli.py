"""Generates tests/golden/example_cli/*: the files the UNMODIFIED reference CLI writes for the bundled example.

  python tests/golden/make_golden_cli.py      (build container only: needs /root/reference; ~40 s on one CPU thread)

Command mirrored: `phertilizer -f example/input/variant_counts.tsv --bin_count_data example/input/binned_read_counts.csv
-s 3 -j 10 --post_process --no-umap --tree_text tree.txt -n cell_clusters.csv -m SNV_clusters.csv --likelihood
likelihood.csv` (README.md:143-149 with --no-umap: umap-learn is not installed here).  The reference's run.main is called
with the Namespace its own get_options() builds.  tests/test_gpu_cli.py feeds the same observations (re-written from
tests/golden/example_inputs.npz with a throw-away first line, which run.py:19-20 drops) to phertilizer_b200.run and
compares the four files byte for byte (likelihoods to 1e-9)."""
import io
import os
import sys
import tempfile
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import example_paths, load_reference  # noqa: E402


def main():
    ref = load_reference()
    run = ref["run"]
    vfile, bfile = example_paths()
    out = os.path.join(HERE, "example_cli")
    os.makedirs(out, exist_ok=True)
    argv = ["-f", vfile, "--bin_count_data", bfile, "-s", "3", "-j", "10", "--post_process", "--no-umap",
            "--tree_text", os.path.join(out, "tree.txt"), "-n", os.path.join(out, "cell_clusters.csv"),
            "-m", os.path.join(out, "SNV_clusters.csv"), "--likelihood", os.path.join(out, "likelihood.csv")]
    saved = sys.argv
    sys.argv = ["phertilizer"] + argv
    try:
        args = run.get_options()
    finally:
        sys.argv = saved
    buf = io.StringIO()
    with redirect_stdout(buf):
        run.main(args)
    open(os.path.join(out, "stdout_tail.txt"), "w").write("\n".join(buf.getvalue().splitlines()[-12:]) + "\n")
    for f in sorted(os.listdir(out)):
        print(f, os.path.getsize(os.path.join(out, f)), "bytes")


if __name__ == "__main__":
    main()
