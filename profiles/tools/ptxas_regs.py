"""Summarise `nvcc -Xptxas=-v` output: kernel template args -> registers / spills / smem."""
import re, subprocess, sys
src = sys.argv[1]
out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                      "-I../../include", "-Xptxas=-v", "-c", src, "-o", "/tmp/_regs.o"],
                     capture_output=True, text=True, cwd="/root/repo/cmad_b200/csrc").stderr
name = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|cmadx::", "", name).split("(")[0]
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        print(f"{name:60s} regs {m.group(1):>3s}  {spill}")
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        spill = f"stack {m.group(1)} spill st {m.group(2)} ld {m.group(3)}"
