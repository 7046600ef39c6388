set -x
timeout -k 10 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2c_pytest.txt
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.txt 2>&1
( time timeout -k 10 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err ) 2> gpurun_out/r2c_bench_time.txt
( time timeout -k 10 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2c_bench_reference.json 2>> gpurun_out/r2c_bench.err ) 2>> gpurun_out/r2c_bench_time.txt
cat gpurun_out/r2c_pytest.txt gpurun_out/r2c_smoke.txt gpurun_out/r2c_bench_time.txt; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench.json')); print(d['value'], d['e2e']['value'], d['fit_step']['golf_mss_loss']['ms_per_step'], d['fit_step']['torch_stft_loss']['ms_per_step']); [print('  ', k['name'], k['kernels'], round(k['ms']*1e3,1)) for k in d['roofline']['kernels']]"
