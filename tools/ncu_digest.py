"""Digest of an `ncu --set full --import-source on` report, small enough to travel back from the GPU box:
   python tools/ncu_digest.py report.ncu-rep [top] > digest.txt
For the first launch of every distinct kernel: duration, registers, issue-active, tensor-pipe-active, executed warp
instructions, DRAM bytes, then the source lines with the most warp-stall samples (ncu_hot_lines.py)."""
import csv, subprocess, sys, os, tempfile
rep = sys.argv[1]
top = sys.argv[2] if len(sys.argv) > 2 else "28"
here = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
keys = [("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__inst_executed.sum", "inst"), ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%")]
seen = set()
for n, r in enumerate(data):
    name = r[col["Kernel Name"]]
    if name in seen:
        continue
    seen.add(name)
    print("=" * 110)
    print(f"[{n}] {name[:100]}")
    print("   " + "  ".join(f"{lab}={r[col[k]]}" for k, lab in keys if k in col))
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
        subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                        f":::{n + 1}"], stdout=f, stderr=subprocess.DEVNULL)
    out = subprocess.run([sys.executable, os.path.join(here, "ncu_hot_lines.py"), f.name, top], capture_output=True,
                         text=True).stdout
    print(out)
    os.unlink(f.name)
