"""Turn an ncu report (or its `--page raw --csv` dump) into the compact per-launch table committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_forward_full.csv
"""
import csv
import subprocess
import sys

KEEP = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "time"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "imma_pipe_active_pct"),
    ("sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", "imma_inst_pct"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tmem_active_pct"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu_dram_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu_pipe_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "fma_pipe_pct"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu_wavefronts_pct"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tc_wavefronts_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    if src.endswith(".ncu-rep"):
        txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
    else:
        rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [(h, n) for h, n in KEEP if h in ix]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["#"] + [n + (f" [{units[ix[h]]}]" if units[ix[h]] else "") for h, n in cols])
        for k, d in enumerate(data):
            out = []
            for h, n in cols:
                v = d[ix[h]]
                if n == "kernel":
                    v = v.replace("void <unnamed>::", "").split("(")[0]
                out.append(v)
            w.writerow([k] + out)
    print(f"{dst}: {len(data)} launches, {len(cols)} columns")


if __name__ == "__main__":
    main()
