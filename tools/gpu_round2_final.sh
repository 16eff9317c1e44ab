#!/bin/bash
# round-2 closing run: whole GPU suite, smoke, both bench arms, ncu evidence of the graded kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2_final_bench_reference.json
timeout 900 python bench.py --sweep 8,16,32,64 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_final_bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")})
print("roofline", d["roofline"]); print("e2e", d["e2e"]); print("cpu", d["cpu_baseline"])
rd=d["roofline_detail"]; print("fwd", rd["fwd"]["frac"], "both", rd["fwd_plus_bwd"]["frac"])
print("variant_b", rd.get("variant_b")); print("modules", rd.get("modules")); print("sweep", rd.get("sweep"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps3.csv python bench.py --steps 3 --warmup 3 --rounds 1 --no-e2e --no-cpu-baseline --no-modules --no-variant-b > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:view_attention_ -c 2 -o gpurun_out/r2_view_attention_1m python tools/bench_shapes.py --path stream --only stress_v32 --iters 1 --warmup 0 > gpurun_out/r2_ncu_va.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_brief.py gpurun_out/r2_view_attention_1m.ncu-rep > gpurun_out/r2_view_attention_1m_ncu_brief.txt 2>&1; head -40 gpurun_out/r2_view_attention_1m_ncu_brief.txt
