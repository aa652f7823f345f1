#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the per-kernel table committed under profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rN_<name>_ncu_summary.csv

One row per profiled launch: duration, DRAM traffic (read + write = the `traffic` of bench.py's roofline
object), DRAM / L2 / pipe utilisation, occupancy, registers, and the top stall reasons.  Runs here (no GPU):
it only reads the report with `ncu -i ... --page raw --csv`.
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "utchmma_bf16_pct"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem_pct"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("smsp__inst_executed.sum", "warp_insts"),
]
STALL_PREFIX2 = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    unit_of = dict(zip(hdr, units))
    out = csv.writer(sys.stdout)
    names = ["kernel", "grid", "block"] + ["%s[%s]" % (short, unit_of.get(full, "")) for full, short in COLS if full in hdr]
    out.writerow(names + ["top_stalls(warps stalled per issue-active cycle)"])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        kname = d.get("Kernel Name", "").split("(")[0].replace("void ", "")
        line = [kname, d.get("Grid Size", ""), d.get("Block Size", "")]
        for full, _ in COLS:
            if full in hdr:
                line.append(d[full])
        stalls = []
        for k in hdr:
            if k.startswith(STALL_PREFIX2) and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[k].replace(",", "")), k[len(STALL_PREFIX2):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        line.append("; ".join("%s %.2f" % (n, v) for v, n in stalls[:4]))
        out.writerow(line)


if __name__ == "__main__":
    main()
