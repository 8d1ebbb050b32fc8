set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
./tools/microbench/coissue_probe 2>&1 | tee gpurun_out/coissue.txt
python tools/stage_times.py --tag base 2>&1 | tail -1 | tee gpurun_out/stage_base.json
for w in 5 7 9 15; do ETHCNN_LIB=$PWD/tools/variants/libethcnn_w$w.so python tools/stage_times.py --tag w$w 2>&1 | tail -1 | tee gpurun_out/stage_w$w.json; done
