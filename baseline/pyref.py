"""The UNMODIFIED Python reference (aweich/esloco v1.0.6) timed on the host cores -- measurement infrastructure only.

`baseline/_ref/` holds the reference installed by `__graft_entry__.build()` in the build container
(`pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>`; git-ignored, shipped to the GPU
box with the snapshot).  Five of its dependencies are not in the image and carry no arithmetic of the hot path
(intervaltree, Bio, plotly, pyfiglet, tqdm_joblib); `baseline/shims/` stands in for them.

bench.py's `cpu_baseline.python_reference` block calls time_python_reference(): the reference's own
run_simulation_iteration (combined_calculations.py:97-126) fanned out with joblib/loky, one iteration per worker, all
host cores -- the mechanism of esloco.py:149-152 (ParallelPbar is a progress-bar wrapper around joblib.Parallel).  Its
count_matches is an O(targets x reads) Python loop (counting.py:19-53), so the named configs cannot run at full size
(config 2: ~13 min per iteration and core, config 3: days); following SURVEY.md section 8(d) it is timed at the
config-2 SHAPE on a 100 Mb genome with 100 and 400 targets, the per-read and per-pair costs a, b of
t = R (a + b T) are solved from the two runs, and the full-size figure is the extrapolation, labelled as such.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SHIMS = os.path.join(HERE, "shims")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "esloco", "combined_calculations.py"))


def _import_reference():
    for p in (SHIMS, REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    # loky workers are fresh interpreters: they find the reference and the stand-ins through PYTHONPATH
    parts = [SHIMS, REF_DIR] + [p for p in os.environ.get("PYTHONPATH", "").split(os.pathsep) if p]
    os.environ["PYTHONPATH"] = os.pathsep.join(dict.fromkeys(parts))
    import esloco.combined_calculations as cc
    return cc


def _params(mean_read_length, coverage):
    return dict(combinations=[(mean_read_length, coverage)], sequenced_data_path=None, sigma=1, min_read_length=1,
                n_barcodes=1, barcode_weights=None, consuming=False, output_path_plots="/tmp", scaling=1.0,
                min_overlap_for_detection=1, no_cov_plots=True)


def run_parallel(cc, n_jobs, n_iterations, params, genome_size, targets, log_file=os.devnull):
    """n_iterations of the reference's run_simulation_iteration on n_jobs loky workers -> (reads placed, seconds)."""
    from joblib import Parallel, delayed
    t0 = time.perf_counter()
    out = Parallel(n_jobs=n_jobs)(delayed(cc.run_simulation_iteration)(i, params, genome_size, targets, None, log_file)
                                  for i in range(n_iterations))
    dt = time.perf_counter() - t0
    reads = 0
    for _results, dists in out:
        d = dists[0]
        reads += sum(int(v) for k, v in d.items() if k not in ("coverage", "mean_read_length", "iteration", "N_bases"))
    return reads, dt


def time_python_reference(n_jobs, full_reads_per_iteration, full_targets, genome_size=100_000_000, coverage=6,
                          mean_read_length=15000, target_counts=(100, 400), insertion_length=5000):
    """Timed sample of the real reference + extrapolation to the full-size workload.  Returns a JSON-able dict."""
    if not available():
        return {"unavailable": "baseline/_ref is missing: run __graft_entry__.build() in the build container"}
    sys.path.insert(0, os.path.dirname(HERE))
    from esloco_b200 import synthetic as S
    cc = _import_reference()
    run_parallel(cc, n_jobs, n_jobs, _params(mean_read_length, 0.02), genome_size,
                 S.random_insertions(genome_size, 4, insertion_length, 1, seed=1))  # start the worker pool, page in
    runs = []
    for T in target_counts:
        targets = S.random_insertions(genome_size, T, insertion_length, 1, seed=2)
        reads, dt = run_parallel(cc, n_jobs, n_jobs, _params(mean_read_length, coverage), genome_size, targets)
        runs.append({"targets": T, "iterations": n_jobs, "reads_placed": reads, "seconds": dt, "reads_per_s": reads / dt,
                     "iterations_per_s": n_jobs / dt})
    # one iteration per worker, all in flight together: seconds ~ per-core time of ONE iteration = R (a + b T)
    (T1, r1), (T2, r2) = [(r["targets"], r["seconds"] / (r["reads_placed"] / r["iterations"])) for r in runs[:2]]
    b = (r2 - r1) / (T2 - T1)
    a = r1 - b * T1
    t_full = full_reads_per_iteration * (a + b * full_targets)
    return {"kind": "reference", "impl": "unmodified aweich/esloco v1.0.6 run_simulation_iteration under joblib/loky (esloco.py:149-152)",
            "cores": n_jobs, "sample": f"config-2 shape at G = {genome_size:.0e}, {coverage}x, one iteration per worker, "
                                      f"targets in {list(target_counts)}", "runs": runs,
            "fit": {"model": "seconds per iteration and core = R * (a + b * T)", "a_us_per_read": a * 1e6, "b_us_per_pair": b * 1e6},
            "extrapolated_full_size": {"reads_per_iteration": full_reads_per_iteration, "targets": full_targets,
                                       "seconds_per_iteration_per_core": t_full, "iterations_per_s": n_jobs / t_full,
                                       "reads_per_s": n_jobs * full_reads_per_iteration / t_full,
                                       "note": "extrapolation, not a measurement: the full-size loop does not finish in bench time"}}


if __name__ == "__main__":
    import json
    n = len(os.sched_getaffinity(0))
    print(json.dumps(time_python_reference(n, 6_184_000, 100), indent=1))
