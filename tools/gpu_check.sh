# One-GPU round check: GPU tests, smoke, the default bench line, the ncu launch list of the same command and one
# `ncu --set full` capture of each hot kernel.  Results land in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
python tools/h2d_probe.py 2>&1 | tail -1 | tee gpurun_out/h2d.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
python bench.py > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; tail -c 2500 gpurun_out/bench_main.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
CONV_PATH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_features_kernel -s 2 -c 1 -o gpurun_out/prof_conv -f python tools/prof_conv_tc.py > gpurun_out/ncu_conv.log 2>&1
CONV_PATH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fc_fused_pair_kernel -s 2 -c 1 -o gpurun_out/prof_fc -f python tools/prof_conv_tc.py > gpurun_out/ncu_fc.log 2>&1
ls -la gpurun_out
