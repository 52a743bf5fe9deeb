#!/usr/bin/env python
"""Static SASS instruction count of each kernel of a cubin, broken down by the ox_stages.cuh function the instruction was
inlined from (nvdisasm -gi line info). Shows what the instruction cache has to stream per step.
    nvdisasm -gi -c X.cubin > x.sass ; python tools/sass_by_function.py x.sass [kernel-substring]"""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "oxide_control_b200", "csrc", "ox_stages.cuh")).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"  (?:static )?OX_HDN? .*?(\w+)\([^;]*\)\s*(?:const\s*)?\{", l)
    if m:
        funcs.append((i, m.group(1)))


def func_of(line):
    name = "?"
    for i, n in funcs:
        if i <= line:
            name = n
        else:
            break
    return name


want = sys.argv[2] if len(sys.argv) > 2 else ""
WRAPPERS = {"step", "forward", "fwd_position", "fwd_velocity", "rk4", "?"}
kernel, cur, stack, in_stack = None, None, [], False
cnt, per = collections.Counter(), collections.Counter()
for l in open(sys.argv[1]):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        kernel = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:   # consecutive //## lines = the inline chain of the next instruction, innermost first
        if not in_stack:
            stack, in_stack = [], True
        stack.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l) and kernel:
        if in_stack:
            in_stack = False
            names = [func_of(ln) for f, ln in reversed(stack) if f == "ox_stages.cuh"]   # outermost first
            stage = next((n for n in names if n not in WRAPPERS), None)
            cur = ("stage:" + stage) if stage else stack[-1][0]
        if want in kernel:
            per[kernel] += 1
            cnt[(kernel, cur)] += 1
for k, v in per.most_common():
    short = re.sub(r"_GLOBAL__N__\w+?_cu_\w+?_\d+", "", k)[:110]
    print(f"== {v} instructions ({v * 16 / 1024:.0f} KB)  {short}")
    for (kk, f), n in sorted(cnt.items(), key=lambda kv: -kv[1]):
        if kk == k and n >= max(200, v // 200):
            print(f"   {n:8d} {100 * n / v:5.1f}%  {f}")
