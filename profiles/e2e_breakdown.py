import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from esloco_b200.session import get_session
from esloco_b200.combined_calculations import run_simulation_batch
from esloco_b200.utils import targets_to_arrays
w = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
B = w["n_barcodes"]; cov = w["coverage"]; NI = min(w["iterations"], 125)
sess = get_session(0); eng = sess.engine
pool = torch.from_numpy(w["pool"].astype(np.uint32)).pin_memory().numpy()
params = dict(n_barcodes=B, barcode_weights=w.get("barcode_weights"), seed=1, scaling=1.0, combinations=[(w["mean_read_length"], cov)], sequenced_data_path=None,
              sigma=1, min_read_length=1, consuming=False, min_overlap_for_detection=1)
def T(f, n=5):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
for _ in range(3):
    sess.invalidate(); run_simulation_batch(range(0,NI), params, w["genome_size"], w["targets"], None, length_pools={0: pool})
def full():
    sess.invalidate(); run_simulation_batch(range(0,NI), params, w["genome_size"], w["targets"], None, length_pools={0: pool})
print("full e2e ms", T(full))
print("targets_to_arrays ms", T(lambda: targets_to_arrays(w["targets"], B)))
keys, ts, te, tb = targets_to_arrays(w["targets"], B)
print("set_targets ms", T(lambda: eng.set_targets(ts, te, tb)))
print("set_length_table ms", T(lambda: eng.set_length_table(pool)))
out = eng.alloc_outputs(NI)
def runonly():
    eng.run_batch(1, 0, 0, NI, cov, out=out)
print("run_batch (index cached) ms", T(runonly))
def dirty_run():
    eng.set_targets(ts, te, tb); eng.run_batch(1, 0, 0, NI, cov, out=out)
print("set_targets + run_batch (index rebuilt) ms", T(dirty_run))
print("alloc_outputs ms", T(lambda: eng.alloc_outputs(NI)))
res = eng.run_batch(1, 0, 0, NI, cov)
print("cpu() ms", T(lambda: res.cpu()))
print("cpu(pinned) ms", T(lambda: res.cpu(pinned=sess.pinned)))
