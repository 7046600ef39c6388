set -x
timeout -k 10 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2y_pytest.txt
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2y_smoke.txt 2>&1
timeout -k 5 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2y_mss.csv python tools/gpu/prof_mss.py > /dev/null 2>&1
python tools/mss_timeline.py gpurun_out/r2y_mss.csv > gpurun_out/r2y_mss_timeline.txt
timeout 120 ./tools/probe/mma_probe > gpurun_out/r2y_mma_probe.txt 2>&1
timeout -k 5 240 python tools/gpu/diag_mss.py 2>&1 | grep "^B=" > gpurun_out/r2y_mss_diag.txt
timeout -k 5 300 python tools/gpu/diag_mss2.py 2>&1 | grep "^B=" >> gpurun_out/r2y_mss_diag.txt
timeout -k 10 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
cat gpurun_out/r2y_pytest.txt gpurun_out/r2y_smoke.txt gpurun_out/r2y_mss_timeline.txt; head -c 600 gpurun_out/r2y_bench.json
