# One-GPU round check: GPU tests, smoke, the default bench line, the ncu launch list of the same command and one
# `ncu --set full` capture of each hot kernel.  Results land in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; tail -c 600 gpurun_out/bench_main.err; tail -c 400 gpurun_out/bench_main.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 700 gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --only-main > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_features_kernel -s 2 -c 1 -o gpurun_out/prof_conv -f python tools/stage_times.py --steps 2 > gpurun_out/ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fc_fused_pair_kernel -s 2 -c 1 -o gpurun_out/prof_fc -f python tools/stage_times.py --steps 2 > gpurun_out/ncu_fc.log 2>&1
ls -la gpurun_out
