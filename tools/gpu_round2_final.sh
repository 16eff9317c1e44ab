#!/bin/bash
# round-2 closing run: whole GPU suite, smoke, both bench arms (every step under a hard timeout)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_final_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_final_bench_reference.json
timeout -s KILL 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_final_bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")})
print("roofline", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
rd=d["roofline_detail"]; print("fwd", rd["fwd"]["frac"], "both", rd["fwd_plus_bwd"]["frac"])
print("variant_b", rd.get("variant_b",{}).get("step_frac")); print("modules", {k:v.get("ms_per_step") for k,v in rd.get("modules",{}).items() if isinstance(v,dict)})
PY
timeout -s KILL 200 python tools/bench_layer.py --out gpurun_out/r2_layer.json 2>&1 | cut -c1-330
timeout -s KILL 200 python tools/profile_module.py 160000 8 64 --no-mod > gpurun_out/r2_module_s3dis_profile.txt 2>&1; head -1 gpurun_out/r2_module_s3dis_profile.txt
timeout -s KILL 200 python tools/profile_module.py 80000 20 128 --no-mod > gpurun_out/r2_module_kitti_profile.txt 2>&1; head -1 gpurun_out/r2_module_kitti_profile.txt
