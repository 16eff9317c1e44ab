#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mlp_layer or bn_act or tc_linear" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2k_pytest.log
timeout 600 python tools/check_tc_gemm.py > gpurun_out/r2_tc_gemm_check2.log 2>&1; echo "tc check rc=$?"; grep "rel_err" gpurun_out/r2_tc_gemm_check2.log | cut -c1-200 | head -40; grep "ms" gpurun_out/r2_tc_gemm_check2.log | cut -c1-260
cp gpurun_out/r2_tc_gemm_check.json gpurun_out/r2_tc_gemm_check2.json
