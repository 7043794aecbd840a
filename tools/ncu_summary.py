"""Summarise an .ncu-rep (no GPU needed): key raw metrics, stall-reason split, hottest source lines.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [launch_index] [top_n]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    r = data[which]
    print("kernel:", r[idx["Kernel Name"]][:110], f"(launch {which} of {len(data)})")
    for k in KEYS:
        if k in idx:
            print(f"  {k:72s} {r[idx[k]]:>18s} {units[idx[k]]}")
    st = [(h, float(r[idx[h]])) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and r[idx[h]] not in ("", "n/a")]
    tot = sum(v for _, v in st) or 1.0
    print("  stall split (warps per issue, share):")
    for h, v in sorted(st, key=lambda kv: -kv[1])[:8]:
        print(f"    {h.split('stalled_')[1].split('_per_issue')[0]:28s} {v:8.2f}  {100 * v / tot:5.1f}%")
    src = run([rep, "--page", "source", "--csv", "--print-source", "sass"])
    # the source page prints one table per launch; split on the "Kernel Name" lines
    blocks, cur = [], []
    for line in src.splitlines():
        if line.startswith('"Kernel Name"'):
            if cur:
                blocks.append(cur)
            cur = []
        cur.append(line)
    if cur:
        blocks.append(cur)
    if which < len(blocks):
        t = list(csv.reader(io.StringIO("\n".join(blocks[which][1:]))))
        h = {n: i for i, n in enumerate(t[0])}
        body = [x for x in t[1:] if len(x) == len(t[0])]
        samp = lambda x: int(x[h["# Samples"]] or 0)
        total = sum(samp(x) for x in body) or 1
        print(f"  hottest SASS lines ({total} samples):")
        stall_cols = [n for n in t[0] if n.startswith("stall_") and "Not Issued" not in n]
        for x in sorted(body, key=lambda x: -samp(x))[:topn]:
            top = sorted(((n, int(x[h[n]] or 0)) for n in stall_cols), key=lambda kv: -kv[1])[:2]
            print(f"    {100 * samp(x) / total:5.1f}%  {x[h['Source']][:86]:86s} {top}")


if __name__ == "__main__":
    main()
