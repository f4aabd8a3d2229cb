#!/usr/bin/env python
"""Command line of phertilizer_b200 -- the flags of the reference CLI (phertilizer/run.py:132-190), unchanged."""
import argparse
import sys

import numpy as np
import pandas as pd

from .fanout import run_many, visible_devices
from .phertilizer import Phertilizer
from .utils import pickle_save


def main(args):
    cols = ["chr", "mutation_id", "cell_id", "base", "var", "total", "copies"]
    # NB: the reference reads with header=None and skips the first line (run.py:19-20); kept for parity
    variant_data = pd.read_table(args.file, sep="\t", header=None, names=cols, skiprows=[0])
    variant_data.drop("copies", inplace=True, axis=1)
    bin_count_data = pd.read_csv(args.bin_count_data) if ".csv" in args.bin_count_data else pd.read_table(args.bin_count_data)
    print("\nInput variant read count data:")
    print(variant_data.head())
    print("\nInput binned read count data:")
    print(bin_count_data.head())

    ctor = dict(alpha=args.alpha, max_copies=args.copies, mode="rd", dim_reduce=not args.no_umap)

    def run_kwargs(i):
        return dict(max_iterations=args.iterations, starts=args.starts, seed=100 * args.seed + i, radius=args.radius,
                    gamma=args.gamma, d=args.min_obs, use_copy_kernel=True, post_process=args.post_process,
                    low_cmb=args.low_cmb, high_cmb=args.high_cmb, nobs_per_cluster=args.nobs_per_cluster)

    devices = visible_devices()
    best_like = -np.inf
    best_tree = best_list = best_loglikes = None
    if args.runs > 1 and len(devices) > 1 and args.embedding is None:
        # independent runs, one worker process per GPU (fanout.py); same winner as the sequential loop below
        print(f"Spreading {args.runs} runs over devices {devices}")
        results, (cell_lookup, mut_lookup) = run_many(variant_data, bin_count_data, ctor,
                                                      [run_kwargs(i) for i in range(args.runs)], devices)
        for i, (tree, tree_list, loglikes, text) in enumerate(results):
            print(f"Run {i + 1}")
            sys.stdout.write(text)
            if tree.norm_loglikelihood > best_like:
                best_like = tree.norm_loglikelihood
                best_tree, best_list, best_loglikes = tree, tree_list, loglikes
            print(f"\nRun {i + 1} complete: Normalized Log Likelihood: {best_like}")
    else:
        ph = Phertilizer(variant_data, bin_count_data, **ctor)
        if args.embedding is not None:
            ph.save_embedding(args.embedding)
        for i in range(args.runs):
            print(f"Run {i + 1}")
            tree, tree_list, loglikes = ph.phertilize(**run_kwargs(i))
            if tree.norm_loglikelihood > best_like:
                best_like = tree.norm_loglikelihood
                best_tree, best_list, best_loglikes = tree, tree_list, loglikes
            print(f"\nRun {i + 1} complete: Normalized Log Likelihood: {best_like}")
        cell_lookup, mut_lookup = ph.get_id_mappings()

    print("\nPhertilizer Tree")
    print(f"Normalized log likelihood: {best_like}")
    print(best_tree)
    print("\nSaving files.....")
    if args.tree_list is not None:
        pickle_save(best_list, args.tree_list)
    if args.tree is not None:
        (best_tree.tree_dot if ".dot" in args.tree else best_tree.tree_png)(args.tree)
    if args.likelihood is not None:
        best_loglikes.to_csv(args.likelihood)
    if args.tree_pickle is not None:
        pickle_save(best_tree, args.tree_pickle)
    pcell, pmut, _, _ = best_tree.generate_results(cell_lookup, mut_lookup)
    if args.pred_cell is not None:
        pcell.to_csv(args.pred_cell, index=False)
    if args.pred_mut is not None:
        pmut.to_csv(args.pred_mut, index=False)
    if args.tree_path is not None:
        best_list.save_the_trees(args.tree_path)
    if args.tree_text is not None:
        best_tree.save_text(args.tree_text)
    print("\nThanks for planting a tree! See you later, friend.")


def get_options(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("-f", "--file", required=True, help="variant/total read counts, unlabeled columns: chr snv cell base var total")
    p.add_argument("--bin_count_data", required=True, help="binned read counts (or embedding) with a `cell` column")
    p.add_argument("-a", "--alpha", type=float, default=0.001, help="per base read error rate")
    p.add_argument("-j", "--iterations", type=int, default=50, help="maximum number of iterations")
    p.add_argument("-s", "--starts", type=int, default=16, help="number of restarts")
    p.add_argument("-d", "--seed", type=int, default=99059, help="seed")
    p.add_argument("--radius", type=float, default=1)
    p.add_argument("-c", "--copies", type=int, default=5, help="max number of copies")
    p.add_argument("--runs", type=int, default=1, help="number of Phertilizer runs")
    p.add_argument("-g", "--gamma", type=float, default=0.95, help="confidence level of the power calculation")
    p.add_argument("--min_obs", type=int, default=10, help="lower bound on the observations of a partition")
    p.add_argument("-m", "--pred-mut", help="output file for mutation clusters")
    p.add_argument("-n", "--pred_cell", help="output file cell clusters")
    p.add_argument("--post_process", action="store_true", help="post-process the inferred tree")
    p.add_argument("--tree", help="output file for png (dot) of Phertilizer tree")
    p.add_argument("--tree_pickle", help="output pickle of Phertilizer tree")
    p.add_argument("--tree_path", help="directory for pngs of all candidate trees")
    p.add_argument("--tree_list", help="pickle file for the ClonalTreeList of all generated trees")
    p.add_argument("--tree_text", help="text file with the edge list of the best clonal tree")
    p.add_argument("--likelihood", help="output file for the likelihoods of the candidate trees")
    p.add_argument("--embedding", type=str, help="where to save the UMAP coordinates")
    p.add_argument("--no-umap", action="store_true", help="do NOT reduce the binned read counts with UMAP")
    p.add_argument("--low_cmb", type=float, default=0.05)
    p.add_argument("--high_cmb", type=float, default=0.15)
    p.add_argument("--nobs_per_cluster", type=int, default=3)
    if argv is None:
        argv = sys.argv[1:]
    return p.parse_args(argv if argv else ["-h"])


def main_cli():
    main(get_options())


if __name__ == "__main__":
    main_cli()
