set -x
timeout -k 10 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2f_pytest.txt
timeout -k 10 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout -k 10 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2f_bench_reference.json 2>> gpurun_out/r2f_bench.err
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2f_ncu_bench.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"osc_knot|osc_flow|noise_fir|ss_response|ss_stitch|ss_solve|room_fir" -s 10 -c 12 -o gpurun_out/r2f_full python tools/prof_step.py 3 > gpurun_out/r2f_ncu_full.log 2>&1
timeout -k 10 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2f_dram.csv python tools/prof_step_dram.py 16 > gpurun_out/r2f_dram.log 2>&1
python tools/dram_per_step.py gpurun_out/r2f_dram.csv 16 8 > gpurun_out/r2f_dram_per_step.txt 2>&1
timeout -k 10 600 ncu --set full --clock-control none -k regex:mss_gemm -s 15 -c 3 -o gpurun_out/r2f_mss_full python tools/gpu/prof_mss.py > /dev/null 2>&1
timeout -k 5 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2f_mss.csv python tools/gpu/prof_mss.py > /dev/null 2>&1
python tools/mss_timeline.py gpurun_out/r2f_mss.csv > gpurun_out/r2f_mss_timeline.txt
timeout -k 5 240 python tools/gpu/diag_mss.py 2>&1 | grep "^B=" > gpurun_out/r2f_mss_diag.txt
timeout -k 5 300 python tools/gpu/diag_mss2.py 2>&1 | grep "^B=" >> gpurun_out/r2f_mss_diag.txt
timeout -k 10 300 python tools/step_events.py 30 > gpurun_out/r2f_step_events.txt 2>&1
cat gpurun_out/r2f_pytest.txt; tail -2 gpurun_out/r2f_bench.err; cat gpurun_out/r2f_dram_per_step.txt; tail -3 gpurun_out/r2f_mss_timeline.txt
