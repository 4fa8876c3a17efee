"""compare two --dump-kernels tables: python tools/cmp_kernels.py old.json new.json"""
import json, sys
a = {k['label']: k for k in json.load(open(sys.argv[1]))['kernels']}
b = {k['label']: k for k in json.load(open(sys.argv[2]))['kernels']}
ta = tb = 0
for l in sorted(set(a) | set(b), key=lambda l: -(a.get(l, b.get(l))['ms_per_step'])):
    x, y = a.get(l), b.get(l)
    xs = x['ms_per_step'] if x else 0; ys = y['ms_per_step'] if y else 0
    ta += xs; tb += ys
    print(f"{l:34s} {(x['avg_ms']*1e3 if x else 0):7.1f}us -> {(y['avg_ms']*1e3 if y else 0):7.1f}us   x{(y or x)['launches_per_step']:<3} {1e3*(ys-xs):+7.1f}us/step")
print(f"sum {ta:.3f} -> {tb:.3f} ms")
