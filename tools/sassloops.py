#!/usr/bin/env python
"""List the loops (backward branches) of a SASS dump that contain a given mnemonic, with their
instruction mix.  usage: sassloops.py dump.sass [MNEMONIC]"""
import re, sys, collections
ins = []
for l in open(sys.argv[1]):
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2)))
want = sys.argv[2] if len(sys.argv) > 2 else 'SHFL.UP'
loops = []
for a, i in ins:
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', i)
    if m and int(m.group(1), 16) < a: loops.append((int(m.group(1), 16), a))
for lo, hi in loops:
    body = [i for a, i in ins if lo <= a <= hi]
    if not any(want in i for i in body): continue
    c = collections.Counter(re.sub(r'^@!?U?P\d\s+', '', i).split()[0].split('.')[0] for i in body)
    fp = c['DFMA'] + c['DMUL'] + c['DADD']
    print(f'loop {lo:#x}-{hi:#x}: {len(body)} instrs, fp64 {fp}, ', dict(c.most_common(16)))
