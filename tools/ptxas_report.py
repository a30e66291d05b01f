"""Registers / spills per kernel from an `nvcc -Xptxas -v` log (stdin or file argument)."""
import re
import sys

text = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers")
for m in pat.finditer(text):
    name = re.search(r"\d+(reflect_kernel\w+?|tsweep_kernel\w+?|fp64_peak\w*?)EEv", m.group(1))
    print(f"{(name.group(1) if name else m.group(1))[:48]:48s} regs {m.group(5):>3s}  stack {m.group(2):>4s}  "
          f"spill st/ld {m.group(3)}/{m.group(4)}")
