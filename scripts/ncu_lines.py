"""Per-source-line summary of an `ncu --page source --csv --print-source=cuda,sass` dump (instructions, stall
samples, shared wavefronts, L1 tag requests per output row).  usage: ncu_lines.py dump.csv n_rows [file-substring]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nrows = float(sys.argv[2])
# the dump holds one block per source file: "File Path", "Function Name", header, lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)


def num(x):
    try:
        return int(x.split("(")[0])
    except Exception:
        return 0


tot_inst = tot_samp = 0
out = []
for b in blocks:
    hdr = None
    for r in b["rows"]:
        if r and r[0] == "Line No":
            hdr = r
            idx = {k: i for i, k in enumerate(hdr)}
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        inst, samp = num(r[idx["Instructions Executed"]]), num(r[idx["# Samples"]])
        if inst == 0 and samp == 0:
            continue
        out.append((b["file"].split("/")[-1], r[0], r[1], inst, samp, num(r[idx["L1 Wavefronts Shared"]]),
                    num(r[idx["L1 Wavefronts Shared Excessive"]]), num(r[idx["L1 Tag Requests Global"]])))
        tot_inst += inst
        tot_samp += samp
print(f"total warp instructions {tot_inst} = {tot_inst / nrows:.0f} per row; samples {tot_samp}")
agg = collections.OrderedDict()
for f, line, src, inst, samp, sh, she, tag in out:
    a = agg.setdefault((f, line), [src, 0, 0, 0, 0, 0])
    a[1] += inst
    a[2] += samp
    a[3] += sh
    a[4] += she
    a[5] += tag
for (f, line), a in agg.items():
    if len(sys.argv) > 3 and sys.argv[3] not in f:
        continue
    print(f"{f[:14]:14} {line:>4} inst/row {a[1] / nrows:7.1f} samp% {100 * a[2] / max(tot_samp, 1):5.1f} shwf/row {a[3] / nrows:6.1f} "
          f"exc {a[4] / nrows:6.1f} tag/row {a[5] / nrows:6.1f} | {a[0].strip()[:80]}")
