"""cProfile of the reference-facing batch call (one step of bench.py's e2e loop): python profiles/e2e_cprofile.py cfg3"""
import cProfile, pstats, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from esloco_b200.session import get_session
from esloco_b200.combined_calculations import run_simulation_batch
w = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
sess = get_session(0)
pool = torch.from_numpy(w["pool"].astype(np.uint32)).pin_memory().numpy()
params = dict(n_barcodes=w["n_barcodes"], barcode_weights=w.get("barcode_weights"), seed=1, scaling=1.0,
              combinations=[(w["mean_read_length"], w["coverage"])], sequenced_data_path=None, sigma=1, min_read_length=1,
              consuming=False, min_overlap_for_detection=w["min_overlap"])
tree = None
if w["mask"]:
    from esloco_b200.utils import build_mask_tree
    tree = build_mask_tree(w["mask"])
NI = min(w["iterations"], 125)
def step():
    sess.invalidate()
    run_simulation_batch(range(0, NI), params, w["genome_size"], w["targets"], tree, length_pools={0: pool}, reuse_host_buffers=True)
for _ in range(3): step()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
