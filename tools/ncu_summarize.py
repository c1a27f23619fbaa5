"""Turn an .ncu-rep (brought back in gpurun_out/) into a small text summary for profiles/.

    python tools/ncu_summarize.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt ["note"]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__cluster_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    lines = [f"# ncu summary of {rep}", f"# {note}", ""]
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        lines.append(f"## kernel: {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
        for h, u, v in zip(hdr, units, row):
            if any(h == k or h.endswith("." + k) for k in KEYS):
                lines.append(f"{h:90s} {v:>18s} {u}")
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    # first kernel block only
    start = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    hdr = src[start]
    data = []
    for r in src[start + 1:]:
        if len(r) != len(hdr):
            break
        data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    agg = sorted(((sum(int(r[ix[s]]) for r in data), s) for s in stalls), reverse=True)
    lines.append(f"## warp stall sampling (all samples = {tot}, {len(data)} SASS instructions)")
    for n, s in agg[:10]:
        lines.append(f"{s:28s} {n:9d}  {100.0 * n / tot:5.1f}%")
    lines.append("")
    lines.append("## top 25 SASS instructions by samples")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:25]:
        st = sorted(((int(r[ix[s]]), s) for s in stalls if int(r[ix[s]]) > 0), reverse=True)[:2]
        lines.append(f"{int(r[ix['# Samples']]):8d} exec={r[ix['Instructions Executed']]:>10s}  {r[ix['Source']].strip()[:70]:70s} {st}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
