N=$1
mkdir -p gpurun_out/scale5
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 50 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/scale5/scale_n1.json 2> gpurun_out/scale5/scale_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 50 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/scale5/scale_n$N.json 2> gpurun_out/scale5/scale_n$N.err
fi
python - <<PY
import json
s=open("gpurun_out/scale5/scale_n$N.json").read()
d=json.loads(s[s.index('{"metric'):])
print(d["n_gpus"], "%.4g"%d["value"], round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()})
PY
