set -e
W=$(mktemp -d); python - "$W" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from oracle import assets
from oracle import ethcnn_oracle as eo
w = sys.argv[1]
assets.materialize(w, "AI")
for name, W, H, nf in (("c1", 768, 512, 1), ("c3", 4928, 3264, 50)):
    base = [eo.synth_frame(W, H, 500 + k) for k in range(5)]
    with open(os.path.join(w, name + ".yuv"), "wb") as f:
        for k in range(nf):
            f.write(base[k % 5].tobytes()); f.write(bytes([128]) * (W * H // 2))
PY
cd $W
B=$GRAFT_REPO_ROOT/hevc-complexity-reduction_b200/bin/video_to_cu_depth
for i in 1 2; do env ETHCNN_TRACE=1 $B c1.yuv 768 512 32 2>&1 | grep -v "^$\|^---\|Predicting"; done
for i in 1 2; do env ETHCNN_TRACE=1 $B c3.yuv 4928 3264 32 2>&1 | grep -v "^$\|^---\|Predicting"; done
nvidia-smi --query-gpu=persistence_mode --format=csv
