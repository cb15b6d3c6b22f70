"""Where one end-to-end step of bench.py goes (host inputs -> device index -> run -> results on the host):
python profiles/e2e_breakdown.py cfg3 [iterations]"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from esloco_b200.session import get_session
from esloco_b200.combined_calculations import run_simulation_batch
from esloco_b200.utils import TargetColumns, targets_to_arrays
w = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
B = w["n_barcodes"]; cov = w["coverage"]; NI = int(sys.argv[2]) if len(sys.argv) > 2 else 125
sess = get_session(0); eng = sess.engine
pool = torch.from_numpy(w["pool"].astype(np.uint32)).pin_memory().numpy()
params = dict(n_barcodes=B, barcode_weights=w.get("barcode_weights"), seed=1, scaling=1.0, combinations=[(w["mean_read_length"], cov)], sequenced_data_path=None,
              sigma=1, min_read_length=1, consuming=False, min_overlap_for_detection=1)
tc = TargetColumns.from_dict(w["targets"], B)
def T(f, n=5):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
def full(mode="moments"):
    sess.invalidate(); run_simulation_batch(range(0,NI), params, w["genome_size"], tc, None, length_pools={0: pool}, results=mode)
for _ in range(3): full()
print("full e2e (moments, columns) ms", T(full))
print("targets_to_arrays(dict) ms", T(lambda: targets_to_arrays(w["targets"], B)))
keys, ts, te, tb = targets_to_arrays(tc, B)
print("3 array copies ms", T(lambda: (ts.copy(), te.copy(), tb.copy())))
print("set_targets ms", T(lambda: eng.set_targets(ts, te, tb)))
def prep():
    eng.set_targets(ts, te, tb); eng.prepare()
print("set_targets + prepare (index build + upload) ms", T(prep))
print("set_length_table ms", T(lambda: eng.set_length_table(pool)))
def cfg():
    sess.invalidate(); sess.configure(w["genome_size"], tc, None, B, w.get("barcode_weights")); sess.set_pool(None, lambda: pool)
print("invalidate + configure + set_pool ms", T(cfg))
out = eng.alloc_outputs(NI)
print("run_batch ms", T(lambda: eng.run_batch(1, 0, 0, NI, cov, out=out)))
print("run_batch + moments ms", T(lambda: eng.moments(eng.run_batch(1, 0, 0, NI, cov, out=out).tallies)))
def warm():
    run_simulation_batch(range(0,NI), params, w["genome_size"], tc, None, length_pools={0: pool}, results="moments")
print("run_simulation_batch moments, inputs resident ms", T(warm))
print("alloc_outputs ms", T(lambda: eng.alloc_outputs(NI)))
