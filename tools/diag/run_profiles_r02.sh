#!/bin/bash
# round-2 profile refresh (run through gpurun): ncu --set full of the first launches of every tokred kernel family in one
# ToMe step (one small report per family: the whole step with sources exceeds gpurun's 64 MiB return limit), the launch list
# of the same forward, and optionally the per-op table (tools/bench_ops.py).  Outputs under gpurun_out/.
cap() {  # family regex, launches
  ncu --profile-from-start off --set full --clock-control none -k regex:"$1" -c $2 -f -o gpurun_out/r02b_$3 \
      python tools/diag/step_launches.py tome_small_patch16_224 0.7 256 > gpurun_out/ncu_$3.log 2>&1 || tail -3 gpurun_out/ncu_$3.log
}
cap attention_kernel 6 attention
cap add_layernorm_kernel 3 add_layernorm
cap tome_match 1 tome_match
cap tome_merge_ln 1 tome_merge_ln
cap patchify_kernel 1 patchify
cap embed_layernorm_kernel 1 embed_layernorm
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_r02b.csv \
    python tools/diag/step_launches.py tome_small_patch16_224 0.7 256 > /dev/null 2>&1
[ "$1" = "ops" ] && { python tools/bench_ops.py --out gpurun_out/ops_r02b.json > gpurun_out/ops_r02b.txt 2>&1; tail -3 gpurun_out/ops_r02b.txt; }
du -sh gpurun_out; ls -la gpurun_out/r02b_*.ncu-rep gpurun_out/launches_step_r02b.csv
