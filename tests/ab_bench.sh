# one-off A/B helper (not a test): bench the library under shredword/lib against build/libtrainer_base.so on the same box
W=${1:-config1_1GB}; S=${2:-3}; WU=${3:-2}
for v in new base new base; do
  if [ $v = base ]; then export SHRED_LIBTRAINER=$PWD/shredword-trainer_b200/build/libtrainer_base.so; else unset SHRED_LIBTRAINER; fi
  python bench.py --workload $W --steps $S --warmup $WU --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v', round(d['value']), round(d['e2e']['value']), d['detail']['bit_exact_vs_golden'], d['detail']['train_s_steps'], 'avg', round(r['avg_launch_us'],2), 'dense', r['dense_launches']['n'], round(r['dense_launches']['avg_launch_us'] or 0,1), round(r['dense_launches']['scan_phase_gbs'] or 0))"
done
