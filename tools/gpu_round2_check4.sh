#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2i_pytest.log; grep -n "^E  " gpurun_out/r2i_pytest.log | head -10
for p in auto stream ring lane; do
  timeout 300 python tools/bench_shapes.py --path $p --out gpurun_out/r2_shapes_$p.json > /dev/null 2>&1; echo "shapes $p rc=$?"
done
python - <<PY
import json
for p in ("auto","lane"):
    try:
        for l in open(f"gpurun_out/r2_shapes_{p}.json"):
            d=json.loads(l); print(p, {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ("shape","fwd_ms","bwd_ms","fwd_frac","bwd_frac","step_frac")})
    except Exception as e: print(p, e)
PY
timeout 300 python tools/bench_rows.py --out gpurun_out/r2_rows.json > gpurun_out/r2_rows.log 2>&1; echo "rows rc=$?"; grep -E "from_dense|merge|projection|zbuffer" gpurun_out/r2_rows.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; head -c 1200 gpurun_out/r2_bench_n1.json; echo
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print(json.dumps(d.get("roofline_detail",{}).get("modules"), indent=0)[:1500])
print(d.get("cpu_baseline"), d.get("e2e",{}).get("value"))
PY
