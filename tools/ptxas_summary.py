"""Summarise `nvcc -Xptxas -v` output: one line per kernel (registers, spills).
usage: nvcc ... -Xptxas -v -c file.cu 2>&1 | python tools/ptxas_summary.py [filter]"""
import re, subprocess, sys
flt = sys.argv[1] if len(sys.argv) > 1 else ""
name = None; spill = ""
for line in sys.stdin:
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(dva::VAParams\)|void dva::", "", name)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m: spill = f"stack={m.group(1)} spill_st={m.group(2)} spill_ld={m.group(3)}"
    m = re.search(r"Used (\d+) registers", line)
    if m and name and flt in name:
        print(f"{name:80s} regs={m.group(1)} {spill}")
