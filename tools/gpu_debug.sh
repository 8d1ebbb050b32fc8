mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout 300 "$@" > gpurun_out/dbg.json 2> gpurun_out/dbg.err; echo "rc=$? $(tail -c 200 gpurun_out/dbg.err | tail -1 | cut -c1-160)"; }
for q in 22 27 32 37; do run "stage_times qp$q" python tools/stage_times.py --qp $q --steps 20; tail -1 gpurun_out/dbg.json | cut -c1-250; done
run "bench default lib" python bench.py --steps 20 --warmup 5 --only-main --no-cpu-baseline
run "bench fc path 2 (single CTA)" env ETHCNN_FC1=fused python bench.py --steps 20 --warmup 5 --only-main --no-cpu-baseline
run "bench x3" python bench.py --steps 200 --warmup 5 --only-main --no-cpu-baseline
