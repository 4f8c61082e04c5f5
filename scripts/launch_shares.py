"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> shares of device time per kernel.
usage: python scripts/launch_shares.py gpurun_out/<list>.csv profiles/<name>   (writes <name>.csv (copy) and <name>_shares.txt)"""
import collections, csv, re, shutil, sys
src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src, errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
cols = rows[hdr]
name_i, val_i, unit_i = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) != len(cols):
        continue
    k = re.sub(r"\(.*", "", r[name_i]).replace("void ", "")
    ns = float(r[val_i].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(r[unit_i], 1)
    tot[k] += ns; cnt[k] += 1
build = {k for k in tot if "row_norms" in k or "to_bf16_tiled" in k or k.startswith("at::")}   # + torch's synthetic-data generator
search_total = sum(v for k, v in tot.items() if k not in build)
shutil.copyfile(src, dst + ".csv")
with open(dst + "_shares.txt", "w") as f:
    f.write(f"# shares of device time per kernel over the ncu launch list {dst}.csv (cold-cache, serialised: compare shares, not absolutes);\n"
            "# corpus build kernels (row_norms, to_bf16_tiled) and bench.py's synthetic-data generation (torch randn) excluded from the total\n")
    for k, v in tot.most_common():
        if k in build:
            continue
        f.write(f"{100 * v / search_total:6.2f}%  {v / 1e6:10.3f} ms  {cnt[k]:4d} launches  {k}\n")
    f.write("# excluded: " + "; ".join(f"{k[:60]} {tot[k] / 1e6:.3f} ms" for k in build) + "\n")
print(open(dst + "_shares.txt").read())
