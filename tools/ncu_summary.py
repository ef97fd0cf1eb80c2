#!/usr/bin/env python
"""ncu report (.ncu-rep) -> text summary of the metrics the roofline discussion uses, one block per captured launch.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_ncu_x.txt "header line" [label1,label2,...]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__cluster_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct']
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "Ghz": 1e9, "Mhz": 1e6}


def main():
    rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3]
    labels = sys.argv[4].split(",") if len(sys.argv) > 4 else []
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# " + header + "\n")
        for k, r in enumerate(data):
            name = r[col["Kernel Name"]]
            label = labels[k] if k < len(labels) else ""
            f.write(f"\n## launch {k}: {label}  [{name[:110]}]\n")
            get = lambda m: float(r[col[m]]) * SCALE.get(units[col[m]], 1) if m in col and r[col[m]] not in ("", "n/a") else None
            t, rd, wr = get('gpu__time_duration.sum'), get('dram__bytes_read.sum'), get('dram__bytes_write.sum')
            ghz, tp = get('sm__cycles_elapsed.avg.per_second'), get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')
            if t and rd is not None and wr is not None:
                f.write(f"# duration {t * 1e3:.4f} ms; DRAM {rd / 1e9:.3f} GB read + {wr / 1e9:.3f} GB written = {(rd + wr) / t / 1e12:.2f} TB/s\n")
            if ghz and tp is not None:
                f.write(f"# tensor pipe active {tp:.1f} % at {ghz / 1e9:.3f} GHz -> {tp / 100 * 8192 * 148 * ghz / 1e12:.0f} TFLOP/s executed (8192 dense 16-bit FLOP/clk/SM)\n")
            for m in WANT:
                if m in col:
                    f.write(f"{m} [ {units[col[m]]} ] = {r[col[m]]}\n")
    print("wrote", out, len(data), "launches")


if __name__ == "__main__":
    main()
