#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/check_tc_gemm.py > gpurun_out/r2_tc_gemm_check6.log 2>&1; echo "tc rc=$?"
grep -E "rel_err" gpurun_out/r2_tc_gemm_check6.log | awk '{print $NF}' | sort -g | tail -2
grep -E "'ms'" gpurun_out/r2_tc_gemm_check6.log
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2g_pytest.log; grep -n "^E  " gpurun_out/r2g_pytest.log | head -10
timeout 300 ncu --set full --clock-control none --import-source on -k regex:va_lane -c 2 -o gpurun_out/r2_lane python tools/bench_shapes.py --path lane --only s3dis_160k_v8_c64,big_1m_v8_c64 --iters 1 --warmup 0 > gpurun_out/r2_ncu_lane.log 2>&1; echo "ncu rc=$?"
python tools/ncu_brief.py gpurun_out/r2_lane.ncu-rep > gpurun_out/r2_lane_ncu_brief.txt 2>&1; cat gpurun_out/r2_lane_ncu_brief.txt
timeout 200 python tools/profile_module.py 160000 8 64 --no-mod > gpurun_out/r2g_module_s3dis.log 2>&1; head -3 gpurun_out/r2g_module_s3dis.log | cut -c1-200
