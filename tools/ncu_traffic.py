"""DRAM traffic per launch from `ncu --set full` reports -> profiles/r02_ncu_traffic.json (read by bench.py's roofline.traffic)
and a markdown table of the headline metrics -> stdout.
    python tools/ncu_traffic.py KEY=path.ncu-rep [KEY=path.ncu-rep ...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
WANT = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct", "launch__registers_per_thread": "regs",
        "launch__grid_size": "grid", "launch__block_size": "block", "sm__cycles_elapsed.max": "cycles",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_wave_pct"}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def read(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    out = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            try:
                out[WANT[h]] = float(v.replace(",", "")) * SCALE.get(u, 1.0)
            except ValueError:
                pass
    return out


def main():
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    print("| key | kernel | time us | DRAM read MB | DRAM write MB | DRAM GB/s | tensor pipe % | warps active % | regs | grid x block |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for arg in sys.argv[1:]:
        key, path = arg.split("=", 1)
        m = read(path)
        data[key] = {"dram_bytes": m.get("rd", 0.0) + m.get("wr", 0.0), "dram_read": m.get("rd"), "dram_write": m.get("wr"),
                     "us": m.get("us"), "tensor_pct": m.get("tensor_pct"), "source": os.path.basename(path), "kernel": m["kernel"]}
        name = m['kernel'].replace('<unnamed>::', '').replace('void ', '').split('(')[0]
        gbs = (m.get('rd', 0) + m.get('wr', 0)) / max(m.get('us', 1e-9), 1e-9) / 1e3
        print(f"| {key} | `{name[:48]}` | {m.get('us', 0):.1f} | {m.get('rd', 0) / 1e6:.1f} | {m.get('wr', 0) / 1e6:.1f} | "
              f"{gbs:.0f} | {m.get('tensor_pct', 0):.1f} | {m.get('warps_pct', 0):.1f} | {m.get('regs', 0):.0f} | "
              f"{m.get('grid', 0):.0f} x {m.get('block', 0):.0f} |")
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
