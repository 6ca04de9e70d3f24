#!/bin/bash
# Round-end measurement batch (GPU box): everything that profiles/rNN_* is made from.  Usage: tools/round_measure.sh r02
R=${1:-r02}
O=gpurun_out
python bench.py > $O/bench_$R.json 2> $O/bench_$R.err
bash tools/esn0_sweep.sh > $O/esn0_sweep_$R.jsonl 2>&1
python tools/modcod_sweep.py --out $O/modcod_sweep_$R.json > $O/modcod_sweep_$R.log 2>&1
python tools/parity_campaign.py --out $O/parity_campaign_$R.json > $O/parity_campaign_$R.log 2>&1; echo "parity campaign rc=$?" >> $O/parity_campaign_$R.log
python tools/ts_bench.py --out $O/ts_bench_$R.json > $O/ts_bench_$R.log 2>&1
python tools/front_bench.py --out $O/front_bench_$R.json > $O/front_bench_$R.log 2>&1
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $O/san_mem_$R.log 2>&1; echo "memcheck rc=$?" >> $O/san_mem_$R.log
compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > $O/san_race_$R.log 2>&1; echo "racecheck rc=$?" >> $O/san_race_$R.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-cpu --pool 4096 > $O/launches_bench_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ldpc_v2_kernel -s 2 -c 1 -o $O/ldpc_$R -f python tools/prof_run.py --pool 8192 --reps 3 > $O/ncu_ldpc_$R.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/front_launches_$R.csv python tools/front_bench.py --out $O/front_bench_ncu_$R.json > /dev/null 2>&1

python tools/dvbs_bench.py --out $O/dvbs_bench_$R.json > $O/dvbs_bench_$R.log 2>&1
python tools/dvbs_inner_bench.py --out $O/dvbs_inner_bench_$R.json > $O/dvbs_inner_bench_$R.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/vit_launches_$R.csv python tools/vit_prof_run.py > $O/vit_launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:acs_kernel -s 5 -c 1 -o $O/vit_acs_$R -f python tools/vit_prof_run.py > $O/ncu_vit_$R.log 2>&1
python tools/pll_bench.py --out $O/pll_bench_$R.json > $O/pll_bench_$R.log 2>&1
python tools/s2_stage_bench.py --out $O/s2_stage_bench_$R.json > $O/s2_stage_bench_$R.log 2>&1
tools/mixed_stream --gpus 1 --out $O/mixed_stream_${R}_1gpu.json > $O/mixed_stream_${R}_1gpu.log 2>&1
tail -n 3 $O/san_mem_$R.log $O/san_race_$R.log $O/parity_campaign_$R.log
