# weak scaling 1 -> 8 GPUs of one box (bench.py's default protocol, as the driver launches it)
NS=${NS:-"2 4 8"}
run() { n=$1; shift; out=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; grep -E "DIST_|MISMATCH" gpurun_out/$out.err | head -4; python -c "
import json
d=json.load(open('gpurun_out/$out.json')); r=d['roofline']
print('$out', d['ms_per_step'], d['value'], d['multi_gpu_bit_identical'], (d.get('e2e') or {}).get('value'), r['stage_ms_per_step'])"; }
python bench.py --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/r2_weak_n1.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_weak_n1.json')); print('n1', d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'))"
for n in $NS; do run $n r2_weak_n$n --steps 20 --warmup 5 $EXTRA; done
