"""Condenses an .ncu-rep (read here, without a GPU) into the per-kernel table committed under profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--json out.json] > profiles/rNN_name.md"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads (spills)"),
    ("smsp__inst_executed_op_local_st.sum", "local stores (spills)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    print(f"# ncu summary of `{rep.split('/')[-1]}` (`ncu --set full --clock-control none`; per launch)\n")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print(f"## {name}\n")
        print("| metric | value |\n|---|---|")
        rec = {"kernel": name}
        for key, label in METRICS:
            if key in col and r[col[key]] != "":
                print(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |")
                rec[key] = [r[col[key]], units[col[key]]]
        stalls = []
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]:
                try:
                    stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("| top stall reasons (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:5]) + " |")
        print()
        out.append(rec)
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
