N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 4 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["gather_check"], d["e2e"]["gather_check"], d["clocks"])
PY
