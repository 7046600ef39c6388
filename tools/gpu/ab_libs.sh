# interleaved A/B of library builds on one box: bash tools/gpu/ab_libs.sh <rounds> lib1.so lib2.so ...  (names under golf_b200/_lib)
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for so in "$@"; do
    GOLF_B200_SO=$PWD/golf_b200/_lib/$so python bench.py --steps 50 --warmup 5 --no-cpu 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$so', 'value %.4g' % d['value'], 'one-at-a-time %.4f ms' % d['ms_per_step_one_at_a_time'], 'e2e %.4g' % d['e2e']['value'], 'e2e_fr %.4g' % d['e2e_frame_rate_f0']['value'], 'ff %.4g' % d['golf_ff']['value'])"
  done
done
