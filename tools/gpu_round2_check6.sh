#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mlp_layer or bn_act or tc_linear" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2k_pytest.log
timeout 300 python tools/bench_layer.py --out gpurun_out/r2_layer.json 2>&1 | tail -8
timeout 200 python tools/profile_module.py 160000 8 64 --no-mod > gpurun_out/r2_module_s3dis_fused.log 2>&1; head -1 gpurun_out/r2_module_s3dis_fused.log; sed -n 5,24p gpurun_out/r2_module_s3dis_fused.log | cut -c1-90,160-230
timeout 200 python tools/profile_module.py 80000 20 128 --no-mod > gpurun_out/r2_module_kitti_fused.log 2>&1; head -1 gpurun_out/r2_module_kitti_fused.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_size.py tests/test_block_down.py -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest2 rc=$?"; tail -5 gpurun_out/r2l_pytest.log
