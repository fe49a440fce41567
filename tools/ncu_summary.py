"""Summarise an `ncu --set full` report into the JSON bench.py reads for roofline.traffic.
    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/r1_v5_ncu_summary.json "capture description" """
import csv, io, json, subprocess, sys

rep, out, desc = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, scale_to=None):
    if name not in col or r[col[name]] in ("", "n/a"):
        return None
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    if scale_to == "MB":
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    if scale_to == "us":
        v *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6}.get(u, 1.0)
    return v


launches = []
for r in data:
    launches.append({
        "kernel": r[col["Kernel Name"]],
        "grid": r[col["Grid Size"]] if "Grid Size" in col else None,
        "block": r[col["Block Size"]] if "Block Size" in col else None,
        "duration_us": get(r, "gpu__time_duration.sum", "us"),
        "dram_read_MB": get(r, "dram__bytes_read.sum", "MB"),
        "dram_write_MB": get(r, "dram__bytes_write.sum", "MB"),
        "tensor_pipe_pct_of_peak_sustained_active": get(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")
        or get(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active") or get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "dram_pct": get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_pct": get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "regs": get(r, "launch__registers_per_thread"),
        "l2_to_sm_read_MB": (get(r, "l1tex__m_xbar2l1tex_read_bytes.sum", "MB")),
    })
json.dump({"capture": desc, "launches": launches}, open(out, "w"), indent=1)
print(f"{len(launches)} launches -> {out}")
for k in launches:
    print(f'{k["kernel"][:70]:70s} {k["duration_us"] or 0:9.1f} us  r {k["dram_read_MB"] or 0:8.1f} MB  w {k["dram_write_MB"] or 0:8.1f} MB  tensor {k["tensor_pipe_pct_of_peak_sustained_active"] or 0:5.1f}%')
