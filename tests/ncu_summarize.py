"""Development aid: markdown table + traffic JSON from an `ncu --set full` report.
    python tests/ncu_summarize.py gpurun_out/prof.ncu-rep [clips] > table.md
Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU)."""
import csv
import io
import json
import re
import subprocess
import sys


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr = r[0]
    return hdr, r[2:]


def short(name):
    name = name.replace("aw::", "").replace("void ", "")
    name = re.sub(r"\(CUtensorMap_st.*", "", name)
    name = re.sub(r"\((const )?(float|double|T1|aw::|unsigned|int|__half).*", "", name)
    return name.replace("(int)", "")


def main():
    rep = sys.argv[1]
    clips = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    hdr, data = rows_of(rep)

    def g(r, k):
        try:
            return float(r[hdr.index(k)])
        except (ValueError, IndexError):
            return float("nan")
    print("| kernel | grid x block | regs | us | DRAM rd MB | DRAM wr MB | DRAM % | issue-active % | tensor pipe % | L2 hit % | warp instr (M) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    traffic = {}
    for r in data:
        name = short(r[hdr.index("Kernel Name")])
        grid = r[hdr.index("Grid Size")].replace(" ", "") if "Grid Size" in hdr else "?"
        block = r[hdr.index("Block Size")].replace(" ", "") if "Block Size" in hdr else "?"
        us = g(r, "gpu__time_duration.sum")
        rd, wr = g(r, "dram__bytes_read.sum") , g(r, "dram__bytes_write.sum")
        print("| `%s` | %s x %s | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.2f |" % (
            name, grid, block, g(r, "launch__registers_per_thread"), us, rd, wr,
            g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            g(r, "lts__t_sector_hit_rate.pct"), g(r, "smsp__inst_executed.sum") / 1e6))
        traffic.setdefault(name, []).append({"us": us, "dram_rd": rd, "dram_wr": wr,
                                             "tensor": g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                                             "issue": g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                             "dram_pct": g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
    sys.stderr.write(json.dumps({"clips": clips, "kernels": traffic}, indent=1))


if __name__ == "__main__":
    main()
