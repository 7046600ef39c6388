set -x
timeout -k 10 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2g_pytest.txt
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2g_smoke.txt 2>&1
timeout -k 10 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
timeout -k 10 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2g_bench_reference.json 2>> gpurun_out/r2g_bench.err
timeout -k 10 300 python tools/step_events.py 30 > gpurun_out/r2g_step_events.txt 2>&1
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -c 1500 --csv --log-file gpurun_out/r2g_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2g_ncu_bench.log 2>&1
cat gpurun_out/r2g_pytest.txt gpurun_out/r2g_smoke.txt; tail -2 gpurun_out/r2g_bench.err; grep -E "CONCURRENT=off|responses\+z|lpc_ss all|room FIR|oscillator" gpurun_out/r2g_step_events.txt
python -c "
import json
d=json.load(open('gpurun_out/r2g_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ms_per_step_one_at_a_time'], d['parity']['rel_rms_max'], d['fit_step'])
r=json.load(open('gpurun_out/r2g_bench_reference.json')); print(r['value'], r['cpu_baseline'])"
