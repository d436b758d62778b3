"""Summarise `ncu --set full` captures (gpurun_out/*.ncu-rep) into profiles/: one text summary per capture and the
machine-readable profiles/ncu_kernel_facts.json that bench.py reads for `roofline.traffic` / `tensor_pipe_pct_ncu`.

    python tools/ncu_facts.py gpurun_out/r2d_ncu_*.ncu-rep
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__cycles_active.avg", "sm__cycles_active.avg",
]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return v * mult


def main():
    facts_path = os.path.join(ROOT, "profiles", "ncu_kernel_facts.json")
    facts = json.load(open(facts_path)) if os.path.exists(facts_path) else {}
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print("no data in", rep)
            continue
        header, units = rows[0], rows[1]
        name = os.path.splitext(os.path.basename(rep))[0]
        short = re.sub(r"^r2[a-z]_", "r2_", name)
        lines = [f"== {name}  (ncu --set full --clock-control none --import-source on; one launch per row; `python tools/ncu_facts.py`)"]
        for r in rows[2:]:
            d = dict(zip(header, r))
            u = dict(zip(header, units))
            kname = d.get("Kernel Name", "").replace("(int)", "").replace("(bool)", "")
            kname = re.sub(r"\(.*", "", kname).replace("void ", "")
            lines.append(f"   Kernel Name                                                   {kname}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
            for k in KEEP:
                if k in d:
                    lines.append(f"   {k:75s} {d[k]:>18s} {u.get(k, '')}")
            if "dram__bytes_read.sum" in d:
                rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
                wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
                dur = float(d["gpu__time_duration.sum"].replace(",", ""))
                dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3}.get(u["gpu__time_duration.sum"], 1e-9)
                lines.append(f"   => DRAM read+write {(rd + wr) / 1e6:.1f} MB in {dur_s * 1e6:.1f} us = {(rd + wr) / dur_s / 1e9:.0f} GB/s")
                key = "gdrn::" + kname.split("gdrn::")[-1].replace("<unnamed>::", "").replace("unnamed>::", "")
                tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
                facts[key] = {"dram_bytes_per_launch": rd + wr, "duration_us": round(dur_s * 1e6, 1),
                              "tensor_pipe_pct": float(tp) if tp not in (None, "", "n/a") else None,
                              "note": f"ncu --set full of one launch (profiles/{short}.txt): DRAM {rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB written"}
            lines.append("")
        txt = "\n".join(lines)
        open(os.path.join(ROOT, "profiles", short + ".txt"), "w").write(txt)
        print(txt)
    json.dump(facts, open(facts_path, "w"), indent=1)


if __name__ == "__main__":
    main()
