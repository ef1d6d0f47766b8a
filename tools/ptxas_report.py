"""Registers / spills / stack per kernel from `build.py --force -v` output (stdin)."""
import re, sys
cur = None
for line in sys.stdin:
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        cur = m.group(1); info = {}
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        info["stack"], info["st"], info["ld"] = m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        print("%-70s regs %3s stack %5s spill st %5s ld %5s" % (cur[:70], m.group(1), info.get("stack", "?"), info.get("st", "?"), info.get("ld", "?")))
        cur = None
