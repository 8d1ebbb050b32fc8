# usage: bash tools/gpu_variants.sh tag1 tag2 ...   (tools/variants/libethcnn_<tag>.so; "default" = the in-tree library)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then python tools/stage_times.py --tag default 2>&1 | tail -1 | tee gpurun_out/stage_default.json
  else ETHCNN_LIB=$PWD/tools/variants/libethcnn_$v.so python tools/stage_times.py --tag $v 2>&1 | tail -1 | tee gpurun_out/stage_$v.json; fi
done
