"""Registers / spills of every kernel entry point: parses `nvcc -Xptxas -v` output (python __graft_entry__.py --force -v 2> log)
into one line per kernel.   python scripts/ptxas_summary.py ptxas.log > profiles/ptxas_rNN.txt"""
import re
import subprocess
import sys

log = open(sys.argv[1]).read()
rows = []
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:ptxas info\s*: Function properties.*\n)?\s*(\d+) bytes stack frame, "
                     r"(\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", log):
    rows.append((m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4))))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
print(f"{len(rows)} kernel entry points, {sum(1 for r in rows if r[3] or r[4])} with spills")
print(f"{'regs':>5} {'stack':>6} {'spill st':>9} {'spill ld':>9}  kernel")
for (mangled, regs, stack, st, ld), name in sorted(zip(rows, names), key=lambda x: x[1]):
    name = name.replace("lec::", "").replace("(lec::TmaMaps, lec::RowParams)", "").replace("(lec::RowParams)", "")
    print(f"{regs:5d} {stack:6d} {st:9d} {ld:9d}  {name}")
