"""Summarise an `ncu --page raw --csv` export: python tools/ncu_summary.py raw.csv out_summary.csv [traffic.json]
Keeps the columns the judge reads; optionally records dram read+write bytes per launch per kernel label
(bench.py's roofline.traffic reads that file; `source` names the summary csv, which is the file committed under
profiles/)."""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
units = rows[1]
keep = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__cycles_active.avg"]
idx = [hdr.index(k) for k in keep if k in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in data:
        w.writerow([r[i] for i in idx])
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
if len(sys.argv) > 3:
    try:
        tr = json.load(open(sys.argv[3]))
    except OSError:
        tr = {}
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    for r in data:
        m = re.search(r"(\w+)<([\d, ]+)>", r[ik].replace("(int)", ""))
        if not m:
            continue
        name, tags = m.group(1), [t.strip() for t in m.group(2).split(",")]
        name = name.replace("_umma_kernel", "_umma")
        label = f"{name}<{tags[0]}>"
        b = float(r[ir]) * SCALE[units[ir]] + float(r[iw]) * SCALE[units[iw]]
        e = tr.setdefault(label, {"bytes": [], "source": sys.argv[2].split("/")[-1]})
        e["bytes"].append(b)
    for e in tr.values():
        if "bytes" in e and isinstance(e["bytes"], list):
            e["dram_bytes_per_launch"] = sum(e["bytes"]) / len(e["bytes"])
    json.dump(tr, open(sys.argv[3], "w"), indent=1, sort_keys=True)
