# compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over the product's CLI on a small clip that pads in both
# directions -- the three hot kernels (conv mma.sync + TMA, fused tcgen05 FC on CTA pairs, gate / gate+map) in one run.
# Results: gpurun_out/sanitizer_<tool>.txt
set -x
mkdir -p gpurun_out
WORK=$(mktemp -d)
python - "$WORK" <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from tools import synth
d = sys.argv[1]
synth.prepare_models(d)
W, H, nf = 456, 264, 3
uv = bytes([128]) * (W * H // 2)
with open(os.path.join(d, "clip.yuv"), "wb") as f:
    for k in range(nf):
        f.write(synth.synth_frame(W, H, 30 + k).tobytes() + uv)
PY
CLI=$PWD/hevc-complexity-reduction_b200/bin/video_to_cu_depth
cd $WORK
$CLI clip.yuv 456 264 32 && cp cu_depth.dat want.dat
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 $CLI clip.yuv 456 264 32 > $OLDPWD/gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit code $? ; output identical: $(cmp -s cu_depth.dat want.dat && echo yes || echo NO)" >> $OLDPWD/gpurun_out/sanitizer_$tool.txt
  tail -4 $OLDPWD/gpurun_out/sanitizer_$tool.txt
done
