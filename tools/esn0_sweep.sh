for e in 10.0 3.1 2.2 1.5 0.4; do
  python bench.py --steps 6 --warmup 3 --no-cpu --esn0 $e 2>/dev/null > gpurun_out/esn0_$e.json
  python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/esn0_$e.json') if l.startswith('{')][-1];print(json.dumps({'esn0_db':$e,'gbit_s':round(d['value'],2),'frames_per_s':round(d['frames_per_s']),'mean_iters':round(d['mean_ldpc_iters'],2),'fer':d['fer'],'e2e_gbit_s':round(d['e2e']['value'],2),'ldpc_ms':round(d['kernel_ms_per_step']['ldpc_v2_kernel'],3),'bch_ms':round(d['kernel_ms_per_step']['bch_kernel'],3)}))"
done
