# A/B of library variants built by profiles/mkvar.sh on one GPU box: VARS="a b" WLS="cfg2 cfg3" bash profiles/ab.sh
set -e
VARS=${VARS:-"base"}; WLS=${WLS:-"cfg2 cfg3 cfg4 chr"}
for v in $VARS; do for w in $WLS; do
  ESL_LIB=esloco_b200/_variants/$v.so python bench.py --workload $w --steps 5 --warmup 3 --batches-per-step 4 --no-cpu-baseline --no-north-star 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('$v $w', round(d['value']/1e9,2), 'G reads/s; bulk ms', round(d['roofline']['kernel_ms_per_launch'],3), 'e2e', round(d['e2e']['value']/1e9,2))"
done; done
