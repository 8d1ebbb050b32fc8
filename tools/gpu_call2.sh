mkdir -p gpurun_out
./tools/microbench/hmma_latency 2>&1 | tee gpurun_out/hmma_latency.txt
for v in NO_STG NO_HMMA NO_ACT NO_STG_NO_ACT; do ETHCNN_LIB=$PWD/tools/variants/libethcnn_$v.so python tools/stage_times.py --tag $v 2>&1 | tail -1 | tee gpurun_out/stage_$v.json; done
