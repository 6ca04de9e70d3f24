python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for o in 1 0; do for pl in 4096 16384; do
echo "occ3=$o pool=$pl"
DVBS2FEC_LDPC_OCC3=$o python bench.py --steps 5 --warmup 3 --no-cpu --pool $pl 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["mean_ldpc_iters"], d["kernel_ms_per_step"])'
done; done
