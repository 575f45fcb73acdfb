#!/bin/bash
# every BASELINE workload through bench.py once (short), one JSON line each -> gpurun_out/workloads_r01.jsonl
out=${1:-gpurun_out/workloads_r01.jsonl}
: > $out
for w in topk_small_kr0.7_b64 tome_small_kr0.7_b256_bf16 evit_base_kr0.5_b128 dyvit_base_kr0.5_b128 dpcknn_small_kr0.25_b256 kmedoids_small_kr0.25_b256 ats_base_kr0.9_b128 sinkhorn_base_kr0.9_b128 patchmerger_base_kr0.9_b128 sit_base_kr0.9_b128; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/err_$w.txt | tail -1 >> $out || echo "{\"workload\": \"$w\", \"failed\": true}" >> $out
done
python - "$out" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    try: d = json.loads(line)
    except Exception: print("BAD", line[:200]); continue
    if d.get("failed"): print(d); continue
    ks = "; ".join(f"{k['kernel']} {k['avg_us']}us {100*k['frac_hbm']:.0f}%" for k in d["kernels"][:4])
    print(f"{d['config']['workload']:32s} {d['value']:9.1f} img/s  e2e {d['e2e']['value']:9.1f}  {d['ms_per_step']:7.2f} ms  tokred {100*d['tokred_share_of_step']:.1f}% | {ks}")
PY
