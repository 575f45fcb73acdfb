for spec in "dpcknn_small_patch16_224 0.25 256 dpcknn" "sinkhorn_base_patch16_224 0.9 128 sinkhorn" "dyvit_base_patch16_224 0.5 256 dyvit"; do
  set -- $spec
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ll2_$4.csv python tools/diag/step_launches.py $1 $2 $3 > /dev/null 2>&1 || echo "ncu failed $4"
done
ls -la gpurun_out/ll2_*.csv
