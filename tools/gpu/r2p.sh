set -x
timeout -k 10 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2p_pytest.txt
timeout -k 10 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
timeout -k 10 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2p_bench_reference.json 2>> gpurun_out/r2p_bench.err
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -c 1500 --csv --log-file gpurun_out/r2p_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2p_ncu_bench.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"osc_knot|osc_flow|noise_fir|ss_response|ss_stitch|ss_solve|ss_tail|room_fir" -s 10 -c 12 -o gpurun_out/r2p_full python tools/prof_step.py 3 > gpurun_out/r2p_ncu_full.log 2>&1
timeout -k 10 300 python tools/step_events.py 30 > gpurun_out/r2p_step_events.txt 2>&1
cat gpurun_out/r2p_pytest.txt; tail -3 gpurun_out/r2p_bench.err; head -c 1500 gpurun_out/r2p_bench.json; tail -3 gpurun_out/r2p_ncu_bench.log; tail -3 gpurun_out/r2p_ncu_full.log; cat gpurun_out/r2p_step_events.txt
