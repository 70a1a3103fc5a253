run() { n=$1; shift; out=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; grep -E "DIST_|MISMATCH" gpurun_out/$out.err | head -4; python -c "
import json
d=json.load(open('gpurun_out/$out.json')); r=d['roofline']
print('$out', d['ms_per_step'], d['value'], d['multi_gpu_bit_identical'], r['stage_ms_per_step'])"; }
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_weak_n1.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_weak_n1.json')); print('n1', d['ms_per_step'], d['value'])"
run 8 r2_weak_n8 --steps 20 --warmup 5 --no-e2e
GB25_BARO_PERSISTENT=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tests/dist_check.py gaussian_islands 64 48 10 5 > gpurun_out/r2_distcheck_n8_fallback.log 2>&1; echo "fallback distcheck rc=$?"; grep -E "DIST_|MISMATCH" gpurun_out/r2_distcheck_n8_fallback.log
run 8 r2_strong_n8 --steps 50 --warmup 5 --no-e2e --scaling strong --no-partition-check
run 4 r2_strong_n4 --steps 50 --warmup 5 --no-e2e --scaling strong --no-partition-check
run 2 r2_strong_n2 --steps 50 --warmup 5 --no-e2e --scaling strong --no-partition-check
run 8 r2_c5_n8 --steps 10 --warmup 3 --no-e2e --no-partition-check --workload tripolar_eighth_degree_tile
