set -x
timeout -k 10 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2u_pytest.txt
timeout -k 10 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2u_dram.csv python tools/prof_step_dram.py 16 > gpurun_out/r2u_dram.log 2>&1
python tools/dram_per_step.py gpurun_out/r2u_dram.csv 16 8 > gpurun_out/r2u_dram_per_step.txt 2>&1
GOLF_BENCH_NOISE=torch timeout -k 10 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2u_dram_torchnoise.csv python tools/prof_step_dram.py 16 > gpurun_out/r2u_dram2.log 2>&1
python tools/dram_per_step.py gpurun_out/r2u_dram_torchnoise.csv 16 8 > gpurun_out/r2u_dram_per_step_torchnoise.txt 2>&1
cat gpurun_out/r2u_pytest.txt gpurun_out/r2u_dram_per_step.txt gpurun_out/r2u_dram_per_step_torchnoise.txt; tail -3 gpurun_out/r2u_dram.log
