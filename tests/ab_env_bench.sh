# one-off A/B helper (not a test): bench variants of the library (environment knobs / build/libtrainer_base.so) on the same box
# usage: bash tests/ab_env_bench.sh <workload> <steps> <warmup> "<label>:<ENV=1 ...>" ...   (label "base" uses build/libtrainer_base.so, labels "alt*" build/libtrainer_<label>.so)
W=$1; S=$2; WU=$3; shift 3
for round in 1 2; do for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  ( if [ "$label" = base ]; then export SHRED_LIBTRAINER=$PWD/shredword-trainer_b200/build/libtrainer_base.so; fi
    case "$label" in alt*) export SHRED_LIBTRAINER=$PWD/shredword-trainer_b200/build/libtrainer_$label.so;; esac
    for e in $envs; do export $e; done
    python bench.py --workload $W --steps $S --warmup $WU --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; c=d['e2e'].get('cold_first_step') or {}
print('$label', round(d['value']), round(d['e2e']['value']), d['detail']['bit_exact_vs_golden'], d['detail']['train_s_steps'], 'avg', round(r['avg_launch_us'],2), 'dense', round(r['dense_launches']['avg_launch_us'] or 0,1), 'phases', [round(r['phase_avg_us'][k],2) for k in ('probe_and_deltas','fold_and_publish')], 'load', round(d['e2e']['load_s_per_step'],3), 'cold', round(c.get('load_s',0),3), round(c.get('train_s',0),3), 'single-CTA', r.get('single_cta_merges_per_step'))" )
done; done
