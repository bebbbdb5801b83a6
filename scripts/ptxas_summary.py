"""Print registers / spills per kernel from nrays_b200/csrc/ptxas.log."""
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "nrays_b200/csrc/ptxas.log"
cur = None
for line in open(path):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        frame = m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        name = re.sub(r"^_ZN3nrb\d+", "", cur)[:40]
        print("%-42s regs %3s  stack %4s  spill st/ld %s/%s" % (name, m.group(1), frame[0], frame[1], frame[2]))
        cur = None
