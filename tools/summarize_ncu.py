"""Summarise one kernel of an .ncu-rep (raw metrics + source-page stall samples) as markdown.
usage: python tools/summarize_ncu.py rep.ncu-rep out.md "title" [kernel_index]"""
import csv, io, subprocess, sys
from collections import Counter
rep, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
kidx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + kidx]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum"]
out = ["# %s\n" % title, "Source: `ncu --set full --clock-control none --import-source on`, file `%s` (kernel #%d).\n" % (rep.split("/")[-1], kidx),
       "## Raw metrics\n", "| metric | unit | value |", "|---|---|---|"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        out.append("| %s | %s | %s |" % (h, u, v[:110]))
out.append("\n## Warp stall reasons (ratio per issue-active)\n\n| reason | value |\n|---|---|")
for h, v in zip(hdr, vals):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and float(v or 0) >= 0.05:
        out.append("| %s | %.2f |" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v)))
if kidx == 0:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    try:
        sh = srows[1]; data = srows[2:]; ix = {h: i for i, h in enumerate(sh)}
        tot = sum(int(r[ix["# Samples"]]) for r in data if len(r) > ix["# Samples"])
        byop = Counter()
        for r in data:
            if len(r) <= ix["# Samples"]: continue
            toks = r[ix["Source"]].split()
            if not toks: continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            byop[op] += int(r[ix["# Samples"]])
        out.append("\n## Stall samples by SASS opcode (%d samples, %d instructions)\n\n| opcode | samples | share |\n|---|---:|---:|" % (tot, len(data)))
        for op, c in byop.most_common(14):
            out.append("| %s | %d | %.1f%% |" % (op, c, 100.0 * c / max(tot, 1)))
    except Exception as e:
        out.append("\n(source page unavailable: %s)" % e)
open(dst, "w").write("\n".join(out) + "\n")
print("wrote", dst)
