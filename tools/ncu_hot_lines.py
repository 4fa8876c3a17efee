"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line:
   python tools/ncu_hot_lines.py file.csv [top]   -> instructions executed + stall samples per (file, line)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = collections.OrderedDict()
tot_i = tot_s = 0
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or cur_file is None: continue
    if r[0].strip() == "": continue            # SASS rows hang below their source line
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def f(name):
        try: return float(r[hdr[name]] or 0)
        except (KeyError, ValueError, IndexError): return 0.0
    ie, sm = f("Instructions Executed"), f("# Samples")
    if ie == 0 and sm == 0: continue
    key = (cur_file, ln)
    a = agg.setdefault(key, [0.0, 0.0, r[1].strip()[:90]])
    a[0] += ie; a[1] += sm
    tot_i += ie; tot_s += sm
print(f"total warp-instructions {tot_i:.0f}, samples {tot_s:.0f}")
for (fn, ln), (ie, sm, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*sm/max(tot_s,1):5.1f}% smp {100*ie/max(tot_i,1):5.1f}% inst  {fn}:{ln:<4} {src}")
byfile = collections.defaultdict(lambda: [0.0, 0.0])
for (fn, ln), (ie, sm, src) in agg.items():
    byfile[fn][0] += ie; byfile[fn][1] += sm
print("-- by file")
for fn, (ie, sm) in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
    print(f"{100*sm/max(tot_s,1):5.1f}% smp {100*ie/max(tot_i,1):5.1f}% inst  {fn}")
