"""Markdown digest of an .ncu-rep (run here, no GPU needed): headline metrics, stall reasons per issued instruction,
stall samples by SASS mnemonic and the hottest SASS lines.   python tools/ncu_digest.py x.ncu-rep > profiles/x.md"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_active.avg"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(path):
    rows = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    print(f"# ncu digest of `{path.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`)\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## {name[:110]}\n")
        print("| metric | value |\n|---|---|")
        for h in WANT:
            if h in hdr:
                print(f"| {h} | {r[hdr.index(h)]} {units[hdr.index(h)]} |")
        st = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("\nStall reasons (warps stalled per issued instruction): " +
              ", ".join(f"{k} {v:.2f}" for v, k in sorted(st, reverse=True)[:8]) + "\n")
    src = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv", "--print-source", "sass"]))))
    his = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if not his:
        return
    hi = his[0]
    h = src[hi]
    ia, isamp, ix = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    data = []
    for r in src[hi + 1:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > isamp and r[0].startswith("0x"):
            data.append((int(r[isamp] or 0), r[ia].strip(), int(r[ix] or 0)))
    tot = sum(d[0] for d in data) or 1
    c = collections.Counter()
    for s_, s, _ in data:
        tk = s.split()
        m = tk[1] if tk[0].startswith("@") and len(tk) > 1 else tk[0]
        c[m.split(".")[0]] += s_
    print("## Stall samples by SASS mnemonic (first captured launch)\n")
    print(f"{tot} samples over {len(data)} SASS instructions: " +
          ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in c.most_common(10)) + "\n")
    print("| samples | % | executed | SASS |\n|---:|---:|---:|---|")
    for s_, s, x in sorted(data, reverse=True)[:12]:
        print(f"| {s_} | {100 * s_ / tot:.1f} | {x} | `{s[:90]}` |")


if __name__ == "__main__":
    main(sys.argv[1])
