import re, subprocess, sys
def parse(path):
    out = {}; name = None
    for line in open(path):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m: name = m.group(1)
        m = re.search(r"Used (\d+) registers", line)
        if m and name: out.setdefault(name, {})["regs"] = int(m.group(1))
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name: out.setdefault(name, {}).update(stack=int(m.group(1)), spill=int(m.group(2)))
    return out
a, b = parse(sys.argv[1]), parse(sys.argv[2])
names = sorted(set(a) | set(b))
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
for n, d in zip(names, dem):
    x, y = a.get(n), b.get(n)
    if x != y: print("CHANGED", d[:150], x, "->", y)
print(len(names), "kernels;", sum(a.get(n) != b.get(n) for n in names), "changed")
