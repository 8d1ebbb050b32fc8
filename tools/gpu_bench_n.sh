# usage: bash tools/gpu_bench_n.sh N   (under gpurun --gpus N): the bench line at N ranks (+ the drop-in through a server sharding over N GPUs)
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m | head -12 > gpurun_out/topo_n$N.txt; nproc >> gpurun_out/topo_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 400 gpurun_out/bench_n$N.err; tail -c 200 gpurun_out/bench_n$N.json
if [ "$2" = wall ]; then ETHCNN_GPUS=$N timeout 600 python tools/cli_wallclock.py --cases config2,config3,config4 --skip-inprocess 2>&1 | tee gpurun_out/cli_wallclock_n$N.txt; fi
