"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table of ONE forward.
usage: python tools/summarize_launches.py gpurun_out/launches.csv profiles/out.md "title" """
import csv, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
seq = []
for r in rows[1:]:
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    seq.append((r[ix["Kernel Name"]], v, r[ix["Grid Size"]], r[ix["Block Size"]]))
starts = [i for i, (n, _, _, _) in enumerate(seq) if "gather_concat" in n]
fw = [x for x in seq[starts[-1]:] if "taco::" in x[0]]
tot = sum(v for _, v, _, _ in fw)
agg = {}
for n, v, g, b in fw:
    k = n.replace("void taco::<unnamed>::", "").replace("taco::<unnamed>::", "").replace("taco::", "").split("(")[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
with open(dst, "w") as f:
    f.write("# %s\n\n" % title)
    f.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none` over `python tools/profile_target.py` "
            "(config 3: N=32, T_in=100, 200 steps, r=5), last forward of the run. Per-launch times are serialised and "
            "cold-cache: compare SHARES.\n\n")
    f.write("Total of %d launches: %.1f us\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n" % (len(fw), tot))
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, c, v, 100 * v / tot))
    f.write("\n## Launch order\n\n| # | kernel | grid | block | us |\n|---:|---|---|---|---:|\n")
    for i, (n, v, g, b) in enumerate(fw):
        k = n.replace("void taco::<unnamed>::", "").replace("taco::<unnamed>::", "").replace("taco::", "").split("(")[0]
        f.write("| %d | `%s` | %s | %s | %.1f |\n" % (i, k, g, b, v))
print("wrote", dst)
