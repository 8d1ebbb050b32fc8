set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
python tools/h2d_probe.py 2>&1 | tail -1 | tee gpurun_out/h2d.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; tail -c 3000 gpurun_out/bench_main.json
ETHCNN_LIB=$PWD/hevc-complexity-reduction_b200/variants/libethcnn_w15.so python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_w15.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/bench_w15.json').read().strip().splitlines()[-1]); print('W15', d['value'], d['stages'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/launches_r1g.csv
