#!/usr/bin/env python
"""Registers / spills of every estep_v2 instantiation from the ptxas logs of the last build."""
import glob, re, sys
for f in sorted(glob.glob('pylda_b200/csrc/build/estep_v2_inst_lk*.ptxas.log')):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '_ZN5pylda8estep_v2ILi(\d+)ELi(\d+)ELi(\d+)E.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
        lk, j, w, stack, ss, sl, regs = m.groups()
        if int(ss) > 0 or len(sys.argv) > 1 and lk == sys.argv[1]:
            print("LK=%s J=%s W=%s regs=%s spill st/ld %s/%s" % (lk, j, w, regs, ss, sl))
