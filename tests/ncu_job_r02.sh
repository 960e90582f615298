set -x
B="python bench.py --steps 1 --warmup 0 --no-side-legs --no-cpu-baseline"
# (1) plain runs first
$B > gpurun_out/r2_ncu_plain10g.json 2> gpurun_out/r2_ncu_plain10g.err || exit 1
python tests/probe_small_train.py config2_10GB 300 > gpurun_out/r2_probe_small.log 2>&1 || exit 1
# (2) full-set captures of single merge launches of the 10 GB configuration (dense -> sparse), bracketed by cudaProfilerStart/Stop
SHRED_PROFILE_MERGES=0,1,100,1000,8000,16000,24000,31000 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_prof_merge10g $B > gpurun_out/r2_ncu_merge10g.log 2>&1
echo "ncu merge rc=$?"
# (3) ingest / count / list-fill kernels of the 10 GB configuration
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_count|k_fill_lists|k_finalize_count|k_tokenize|k_symbolize|k_hist_words' -c 24 -o gpurun_out/r2_prof_ingest10g python tests/probe_small_train.py config2_10GB 300 > gpurun_out/r2_ncu_ingest10g.log 2>&1
echo "ncu ingest rc=$?"
# (4) launch list (durations) of the first 3000 launches of a 10 GB step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches10g_first3000.csv $B > gpurun_out/r2_ncu_launches10g.log 2>&1
echo "ncu launches rc=$?"
