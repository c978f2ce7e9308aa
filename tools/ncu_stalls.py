"""Per-instruction stall summary of one kernel from `ncu -i X.ncu-rep --page source --csv`.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv
    python tools/ncu_stalls.py /tmp/src.csv [n_top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# the file holds one block per profiled launch: "Kernel Name" line, header, instructions
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
b = blocks[0]
hdr, data = b["hdr"], b["data"]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
texe = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(b["name"][:100])
print("samples", tot, "warp instructions executed", texe)
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stallcols}
print({k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:ntop]:
    st = {h: int(r[ix[h]]) for h in stallcols if int(r[ix[h]]) > 0}
    best = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(r[ix["Address"]][-4:], r[ix["Source"]].strip()[:52].ljust(52), r[ix["# Samples"]].rjust(6),
          "%4.1f%%" % (100 * int(r[ix["# Samples"]]) / tot), best)
