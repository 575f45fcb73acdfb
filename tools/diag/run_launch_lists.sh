set -e
python -m pytest tests -m gpu -x -q -k "gather_rows or ats" > gpurun_out/pytest_s6.log 2>&1; tail -2 gpurun_out/pytest_s6.log
for spec in "tome_small_patch16_224 0.7 256 tome" "dpcknn_small_patch16_224 0.25 256 dpcknn" "evit_base_patch16_224 0.5 256 evit" "sinkhorn_base_patch16_224 0.9 128 sinkhorn" "ats_base_patch16_224 0.9 128 ats"; do
  set -- $spec
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ll_$4.csv python tools/diag/step_launches.py $1 $2 $3 > /dev/null 2>&1 || echo "ncu failed $4"
done
ls -la gpurun_out/ll_*.csv
