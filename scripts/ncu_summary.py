"""Key metrics of every kernel instance in an ncu report (read here, no GPU): python scripts/ncu_summary.py <rep> [...]
Prints one markdown table per report: duration, DRAM bytes and throughput, L2 hit rate, tensor pipe, shared-memory data
pipe, registers, grid."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/smem data pipe (LSU) %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe: tensor-core operand reads %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
]


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


for rep in sys.argv[1:]:
    hdr, units, rows = load(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"\n### `{rep.split('/')[-1]}`\n")
    print("| # | kernel | " + " | ".join(lbl for _, lbl in KEYS) + " |")
    print("|---|---|" + "---:|" * len(KEYS))
    for n, r in enumerate(rows):
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for k, _ in KEYS:
            if k in idx:
                v, u = r[idx[k]], units[idx[k]]
                try:
                    f = float(v)
                    v = f"{f:.3f}" if abs(f) < 1000 else f"{f:.0f}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("-")
        print(f"| {n} | `{name}` | " + " | ".join(cells) + " |")
