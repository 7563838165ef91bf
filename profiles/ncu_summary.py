#!/usr/bin/env python
"""Condense an ncu report (`ncu --set full`) into one line per launch: duration, DRAM bytes and throughput, issue
activity, registers, shared-memory bank conflicts.  Usage: ncu_summary.py report.ncu-rep > profiles/ncu/<name>.txt"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "inst"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    cols = [(hdr.index(m), lab) for m, lab in WANT if m in hdr]
    print("# %s" % path)
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        parts = ["%s=%s%s" % (lab, r[i].split(".")[0] if "." in r[i] and len(r[i]) > 9 else r[i],
                              "" if units[i] in ("", "inst", "register/thread") else units[i]) for i, lab in cols]
        print("%-44s %s" % (name[:44], " ".join(parts)))


if __name__ == "__main__":
    main(sys.argv[1])
