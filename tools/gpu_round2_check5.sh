#!/bin/bash
mkdir -p gpurun_out
for p in auto lane; do
  timeout 300 python tools/bench_shapes.py --path $p --out gpurun_out/r2_shapes_$p.json > /dev/null 2>&1; echo "shapes $p rc=$?"
done
python - <<PY
import json
for p in ("auto",):
    for l in open(f"gpurun_out/r2_shapes_{p}.json"):
        d=json.loads(l); print(p, {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ("shape","fwd_ms","bwd_ms","fwd_frac","bwd_frac","step_frac")})
PY
for n in 0 1; do
  DVA_TC_NARROW=$n timeout 200 python tools/profile_module.py 160000 8 64 --no-mod > gpurun_out/r2_module_s3dis_narrow$n.log 2>&1; echo "narrow=$n"; head -1 gpurun_out/r2_module_s3dis_narrow$n.log | cut -c1-200
  DVA_TC_NARROW=$n timeout 200 python tools/profile_module.py 80000 20 128 --no-mod > gpurun_out/r2_module_kitti_narrow$n.log 2>&1; head -1 gpurun_out/r2_module_kitti_narrow$n.log | cut -c1-200
done
sed -n 5,30p gpurun_out/r2_module_s3dis_narrow1.log | cut -c1-90,160-230
DVA_TC_NARROW=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_size.py -m gpu -q -k "pool or linear or bn" > gpurun_out/r2j_pytest_narrow.log 2>&1; echo "narrow pytest rc=$?"; tail -3 gpurun_out/r2j_pytest_narrow.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:va_lane -c 2 -o gpurun_out/r2_lane_final python tools/bench_shapes.py --path auto --only s3dis_160k_v8_c64,big_1m_v8_c64 --iters 1 --warmup 0 > gpurun_out/r2_ncu_lane3.log 2>&1; echo "ncu rc=$?"
python tools/ncu_brief.py gpurun_out/r2_lane_final.ncu-rep > gpurun_out/r2_lane_final_ncu_brief.txt 2>&1; head -24 gpurun_out/r2_lane_final_ncu_brief.txt
