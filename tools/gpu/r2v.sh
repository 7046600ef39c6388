set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  timeout -k 10 300 $TR --nproc-per-node $n --master-port $((29600+n)) tools/fit_step.py 20 ss 2>/dev/null | grep '^{' > gpurun_out/r2v_fit_n$n.json
done
timeout -k 10 300 python tools/fit_step.py 20 ss 2>/dev/null | grep '^{' > gpurun_out/r2v_fit_n1.json
timeout -k 10 300 $TR --nproc-per-node 8 --master-port 29650 tools/fit_step.py 20 ff 2>/dev/null | grep '^{' > gpurun_out/r2v_fit_ff_n8.json
for n in 8 4 2; do
  timeout -k 10 400 $TR --nproc-per-node $n --master-port $((29700+n)) bench.py --gpus $n --steps 50 --warmup 5 --no-cpu > gpurun_out/r2v_bench_n$n.json 2> gpurun_out/r2v_bench_n$n.err
done
cat gpurun_out/r2v_fit_n*.json | cut -c1-700
for n in 8 4 2; do python -c "
import json; d=json.load(open('gpurun_out/r2v_bench_n$n.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('copy_only_ms_per_step'))"; done
