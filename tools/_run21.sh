NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_run.py 65536 4 > /dev/null && $NCU -k regex:k_step -s 3 -c 1 -f -o gpurun_out/r2r_fl python tools/prof_run.py 65536 4 > gpurun_out/r2r_ncu_fl.log 2>&1; tail -1 gpurun_out/r2r_ncu_fl.log
$NCU -k regex:k_policy_mlp -s 2 -c 1 -f -o gpurun_out/r2r_pol python tools/prof_run.py 65536 4 > gpurun_out/r2r_ncu_pol.log 2>&1; tail -1 gpurun_out/r2r_ncu_pol.log
$NCU -k regex:k_step -s 2 -c 1 -f -o gpurun_out/r2r_w4 python tools/prof_run.py 16384 3 w4_stairs > gpurun_out/r2r_ncu_w4.log 2>&1; tail -1 gpurun_out/r2r_ncu_w4.log
$NCU -k regex:k_step -s 2 -c 1 -f -o gpurun_out/r2r_hum python tools/prof_run.py 32768 3 humanoid_slope > gpurun_out/r2r_ncu_hum.log 2>&1; tail -1 gpurun_out/r2r_ncu_hum.log
cp cosim_b200/csrc/_build/engine.cu.o gpurun_out/r2r_engine.cu.o
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --steady-steps 0 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --steady-steps 0 > gpurun_out/r2r_launch_bench.log 2>&1
python bench.py > gpurun_out/r2r_bench_headline.json 2> gpurun_out/r2r_bench_headline.err; tail -c 600 gpurun_out/r2r_bench_headline.json
python bench.py --policy zero --no-cpu-baseline > gpurun_out/r2r_bench_headline_zero.json 2>> gpurun_out/r2r_bench_headline.err
python bench.py --config w4_stairs --steps 10 --warmup 3 > gpurun_out/r2r_bench_w4.json 2> gpurun_out/r2r_bench_w4.err
python bench.py --config humanoid_slope --steps 10 --warmup 3 > gpurun_out/r2r_bench_hum.json 2> gpurun_out/r2r_bench_hum.err
python bench.py --config light_flat --no-cpu-baseline > gpurun_out/r2r_bench_light.json 2> gpurun_out/r2r_bench_light.err
python bench.py --config flamingo_rocky_norand --no-cpu-baseline > gpurun_out/r2r_bench_norand.json 2> gpurun_out/r2r_bench_norand.err
for f in gpurun_out/r2r_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('steady_state',{}).get('value'), d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))
except Exception as e: print('ERR', e)
"; done
