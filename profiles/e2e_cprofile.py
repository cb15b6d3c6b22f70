"""cProfile of the reference-facing batch call (one step of bench.py's e2e loop): python profiles/e2e_cprofile.py cfg3 125"""
import cProfile, pstats, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from esloco_b200.session import get_session
from esloco_b200.combined_calculations import run_simulation_batch
from esloco_b200.utils import TargetColumns
w = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
NI = int(sys.argv[2]) if len(sys.argv) > 2 else 125
B = w["n_barcodes"]
sess = get_session(0)
pool = torch.from_numpy(w["pool"].astype(np.uint32)).pin_memory().numpy()
params = dict(n_barcodes=B, barcode_weights=w.get("barcode_weights"), seed=1, scaling=1.0, combinations=[(w["mean_read_length"], w["coverage"])],
              sequenced_data_path=None, sigma=1, min_read_length=1, consuming=False, min_overlap_for_detection=1)
tc = TargetColumns.from_dict(w["targets"], B)
def step():
    sess.invalidate()
    res, _ = run_simulation_batch(range(0, NI), params, w["genome_size"], tc, None, length_pools={0: pool}, results="moments", to_host=False)
    torch.cuda.synchronize()
for _ in range(3): step()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
