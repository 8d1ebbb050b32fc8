mkdir -p gpurun_out
for w in 1 3 7; do ETHCNN_LIB=$PWD/tools/variants/libethcnn_Tw$w.so python tools/stage_times.py --tag Tw$w --steps 5 2>&1 | tail -1 | tee gpurun_out/stage_Tw$w.json; done
timeout 1500 python -m pytest tests/test_gpu_real_content.py tests/test_gpu_decision_map.py tests/test_gpu_hm_live.py "tests/test_gpu_parity.py::test_resident_server_serves_the_drop_in" tests/test_gpu_ldp.py -m gpu -x -q -s 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_new.txt
