#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mlp_layer or bn_act or tc_linear" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2k_pytest.log
timeout 600 python tools/check_tc_gemm.py > gpurun_out/r2_tc_gemm_check2.log 2>&1; echo "tc check rc=$?"; grep "ms" gpurun_out/r2_tc_gemm_check2.log | cut -c1-260
cp gpurun_out/r2_tc_gemm_check.json gpurun_out/r2_tc_gemm_check2.json
timeout 300 python tools/bench_layer.py --out gpurun_out/r2_layer.json 2>&1 | tail -8
timeout 200 python tools/profile_module.py 160000 8 64 --no-mod > gpurun_out/r2_module_s3dis_fused.log 2>&1; head -1 gpurun_out/r2_module_s3dis_fused.log; sed -n 5,20p gpurun_out/r2_module_s3dis_fused.log | cut -c1-90,160-230
timeout 200 python tools/profile_module.py 80000 20 128 --no-mod > gpurun_out/r2_module_kitti_fused.log 2>&1; head -1 gpurun_out/r2_module_kitti_fused.log
timeout 400 python bench.py --no-e2e --no-cpu-baseline --no-modules --steps 10 --warmup 3 --rounds 3 > gpurun_out/r2_bench_variantb.json 2> gpurun_out/r2_bench_variantb.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_variantb.json')); print(d['value'], d['roofline']['frac']); print(d['roofline_detail'].get('variant_b'))"
