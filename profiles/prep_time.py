import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from esloco_b200.session import get_session
from esloco_b200.utils import TargetColumns, targets_to_arrays
w = bench.make_workload("cfg3")
sess = get_session(0); eng = sess.engine
tc = TargetColumns.from_dict(w["targets"], 24)
keys, ts, te, tb = targets_to_arrays(tc, 24)
eng.set_genome(w["genome_size"]); 
from esloco_b200.read_operations import barcode_cdf, barcode_probabilities
eng.set_barcode_cdf(barcode_cdf(barcode_probabilities(24, w["barcode_weights"])))
eng.set_blocked(); eng.set_length_table(w["pool"])
for i in range(3):
    eng.set_targets(ts, te, tb); torch.cuda.synchronize(); t=time.perf_counter(); eng.prepare(); torch.cuda.synchronize(); print("prepare ms", (time.perf_counter()-t)*1e3, file=sys.stderr)
