#!/bin/bash
# usage: scripts/stage_times.sh <tag> [defines]   -> rebuild with defines, run a short bench, print stage times
tag=$1; shift
GB25_NVCC_DEFINES="$*" python -c "import gb25_b200.build as b; b.build_cuda(force=True)" || exit 1
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline ${GB25_WORKLOAD:+--workload $GB25_WORKLOAD} > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { tail -5 gpurun_out/bench_$tag.err; exit 1; }
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print('$tag', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"
