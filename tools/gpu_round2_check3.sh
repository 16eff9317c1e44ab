#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2h_pytest.log; grep -n "^E  " gpurun_out/r2h_pytest.log | head -10
for p in auto lane; do
  timeout 200 python tools/bench_shapes.py --path $p --only s3dis_160k_v8_c64,s3dis_160k_v8_c64_bf16,pyramid_160k_v8_c32,big_1m_v8_c64,big_1m_v8_c64_bf16,stress_v8,sphere_40k_v8_c64 --out gpurun_out/r2b_shapes_$p.json > /dev/null 2>&1; echo "shapes $p rc=$?"
done
python - <<PY
import json
for p in ("auto","lane"):
    try:
        for l in open(f"gpurun_out/r2b_shapes_{p}.json"):
            d=json.loads(l); print(p, {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ("shape","name","fwd_ms","bwd_ms","fwd_frac","bwd_frac","step_frac")})
    except Exception as e: print(p, e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:va_lane -c 4 -o gpurun_out/r2_lane2 python tools/bench_shapes.py --path lane --only s3dis_160k_v8_c64,big_1m_v8_c64 --iters 1 --warmup 0 > gpurun_out/r2_ncu_lane2.log 2>&1; echo "ncu rc=$?"
python tools/ncu_brief.py gpurun_out/r2_lane2.ncu-rep > gpurun_out/r2_lane2_ncu_brief.txt 2>&1; cat gpurun_out/r2_lane2_ncu_brief.txt | head -60
