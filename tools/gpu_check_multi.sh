# Run on a box with N >= 2 GPUs (gpurun --gpus N): peer-gather tests, then bench.py at N ranks with the peer-memory
# gather and, for comparison, with the NCCL gather.
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py::test_in_library_multi_gpu_matches_single -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_multi.txt
for mode in "" "--nccl-gather"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 4 $mode > gpurun_out/bench_n${N}${mode}.json 2> gpurun_out/bench_n${N}${mode}.err
  tail -c 1200 gpurun_out/bench_n${N}${mode}.err; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n${N}${mode}.json").read().strip().splitlines()[-1])
print("N=${N} ${mode}", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["sharding"][:60], d["config"]["gather_check"], d["stages"]["gate"])
PY
done
