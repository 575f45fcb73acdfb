#!/bin/bash
# final round-2 artefacts (run through gpurun): full GPU suite, default bench line (+ workloads array), per-op table
python -m pytest tests -m gpu -q > gpurun_out/pytest_r02.log 2>&1; tail -3 gpurun_out/pytest_r02.log
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -c 300 gpurun_out/bench_r02.err
python tools/bench_ops.py --out gpurun_out/ops_r02.json > gpurun_out/ops_r02.txt 2>&1; tail -2 gpurun_out/ops_r02.txt
