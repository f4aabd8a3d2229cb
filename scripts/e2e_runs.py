"""`--runs R` end to end on N GPUs (BASELINE.json configs[3] pattern): wall time of R independent phertilize() runs on
synthetic data, sequential on one device (the reference's loop, run.py:55-74) or spread by phertilizer_b200.fanout.

  python scripts/e2e_runs.py C2 --runs 8 --starts 4 --iterations 10 --devices 8
Prints one JSON line; wall times include the workers' data builds (one per device)."""
import argparse, io, json, sys, time
from contextlib import redirect_stdout

sys.path.insert(0, ".")

ap = argparse.ArgumentParser()
ap.add_argument("config", nargs="?", default="C2")
ap.add_argument("--runs", type=int, default=8)
ap.add_argument("--starts", type=int, default=4)
ap.add_argument("--iterations", type=int, default=10)
ap.add_argument("--devices", type=int, default=1)
ap.add_argument("--seed", type=int, default=99059)
args = ap.parse_args()

if __name__ == "__main__":
    import torch

    from phertilizer_b200.fanout import best_run, run_many
    from phertilizer_b200.phertilizer import Phertilizer
    from phertilizer_b200.synth import make_config, to_dataframes

    s = make_config(args.config, device="cuda" if torch.cuda.is_available() else "cpu")
    vc, bc = to_dataframes(s)
    ctor = dict(alpha=0.001, max_copies=5, mode="rd", dim_reduce=False)
    kws = [dict(max_iterations=args.iterations, starts=args.starts, seed=100 * args.seed + i, radius=1, gamma=0.95, d=10,
                use_copy_kernel=True, post_process=True) for i in range(args.runs)]
    del s
    torch.cuda.empty_cache()
    t0 = time.time()
    if args.devices == 1:
        ph = Phertilizer(vc, bc, **ctor)
        res = []
        for kw in kws:
            with redirect_stdout(io.StringIO()):
                res.append(ph.phertilize(**kw))
        norms = [r[0].norm_loglikelihood for r in res]
        best = max(range(len(norms)), key=lambda i: (norms[i], -i))
    else:
        results, _ = run_many(vc, bc, ctor, kws, list(range(args.devices)))
        norms = [r[0].norm_loglikelihood for r in results]
        best = best_run(results)
    wall = time.time() - t0
    print(json.dumps({"config": args.config, "runs": args.runs, "starts": args.starts, "iterations": args.iterations,
                      "devices": args.devices, "wall_s": round(wall, 2), "best_run": int(best),
                      "norm_loglikelihoods": [float(x) for x in norms]}))
