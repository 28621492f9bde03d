#!/usr/bin/env python
"""Static SASS instruction count per source line of a cubin built with -lineinfo:
   nvdisasm -g -c x.cubin > x.sass; python sass_lines.py x.sass [top]      (development tool)"""
import collections, re, sys
cur = None
by_line, by_op = collections.Counter(), collections.Counter()
inl = []
for l in open(sys.argv[1]):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m:
        by_line[cur] += 1
        by_op[m.group(2).split(".")[0]] += 1
tot = sum(by_line.values())
print("static instructions:", tot)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
import os
for (f, n), v in by_line.most_common(top):
    try: text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "../../bbm_b200/csrc", f)).read().split("\n")[n-1].strip()[:110]
    except Exception: text = ""
    print("%-22s %5d %5d %5.1f%%  %s" % (f, n, v, 100.0*v/tot, text))
print(by_op.most_common(40))
