"""ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) -> markdown table + per-kernel traffic JSON.
    python tools/ncu_table.py gpurun_out/r01b_train_raw.csv profiles/r01b_ncu_train_step_full.md profiles/r01b_traffic.json
"""
import csv
import json
import re
import sys

COLS = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (hmma cycles active)"),
        ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "bf16 tensor ops % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue %")]


def short(name):
    m = re.search(r"(\w+_kernel(?:<[^>(]*>)?)", name)
    return m.group(1) if m else name


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(src, md, js):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k) if k in hdr else None, t) for k, t in COLS]
    out = ["| " + " | ".join(t for _, t in COLS) + " |", "|" + "---|" * len(COLS)]
    traffic = {}
    for r in rows[2:]:
        cells = []
        for i, t in idx:
            if i is None:
                cells.append("n/a")
            elif t == "kernel":
                cells.append(short(r[i]))
            else:
                cells.append((r[i] + " " + units[i]).strip())
        out.append("| " + " | ".join(cells) + " |")
        k = short(r[hdr.index("Kernel Name")])
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        traffic.setdefault(k, []).append({"dram_read_bytes": to_bytes(r[ir], units[ir]),
                                          "dram_write_bytes": to_bytes(r[iw], units[iw]),
                                          "duration": r[it] + " " + units[it], "grid": r[hdr.index("Grid Size")]})
    open(md, "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(js, "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main(*sys.argv[1:4])
