#!/bin/bash
# One GPU-box session of round 2: parity tests, the tcgen05 GEMM check, the short-segment shape sweep per
# implementation of the fused pair, and an ncu capture of the tcgen05 kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2f_pytest.log; grep -n "^E  " gpurun_out/r2f_pytest.log | head -10
timeout 300 python tools/check_tc_gemm.py > gpurun_out/r2_tc_gemm_check5.log 2>&1; echo "tc rc=$?"
grep -E "dW" gpurun_out/r2_tc_gemm_check5.log; grep -E "'ms'" gpurun_out/r2_tc_gemm_check5.log | grep -v dW | head
for p in auto stream ring lane; do
  timeout 200 python tools/bench_shapes.py --path $p --only s3dis_160k_v8_c64,s3dis_160k_v8_c64_bf16,pyramid_160k_v8_c32,big_1m_v8_c64,big_1m_v8_c64_bf16,stress_v8,kitti_80k_v20_c128,sphere_40k_v8_c64 --out gpurun_out/r2_shapes_$p.json > /dev/null 2>&1; echo "shapes $p rc=$?"
done
python - <<PY
import json
for p in ("auto","stream","ring","lane"):
    try:
        for l in open(f"gpurun_out/r2_shapes_{p}.json"):
            d=json.loads(l); print(p, d.get("name"), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ("fwd_ms","bwd_ms","fwd_frac","bwd_frac","step_frac")})
    except Exception as e: print(p, e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_ -c 6 -o gpurun_out/r2_tc_gemm python tools/run_tc_once.py > gpurun_out/r2_ncu_tc.log 2>&1; echo "ncu rc=$?"
python tools/ncu_brief.py gpurun_out/r2_tc_gemm.ncu-rep > gpurun_out/r2_tc_gemm_ncu_brief.txt 2>&1; head -70 gpurun_out/r2_tc_gemm_ncu_brief.txt
