"""Aggregate the ncu source page of one kernel by SASS opcode: executed warp-instructions and stall samples.
usage: python tools/ncu_src.py rep kernel_regex [top_lines]"""
import csv, subprocess, sys, io, collections, re
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}; blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
b = blocks[0]
h = {n: i for i, n in enumerate(b["hdr"])}
agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for r in b["rows"]:
    src = r[h["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)", src)
    op = m.group(2).split(".")[0] if m else src[:10]
    ex = int(r[h["Instructions Executed"]] or 0); sm = int(r[h["# Samples"]] or 0)
    agg[op][0] += ex; agg[op][1] += sm; tot_i += ex; tot_s += sm
print(b["name"][:100], "total warp-instr", tot_i, "samples", tot_s)
for op, (ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"  {op:10s} exec {ex:9d} ({100*ex/tot_i:5.1f}%)  samples {sm:6d} ({100*sm/max(tot_s,1):5.1f}%)")
if top:
    print("hottest lines by samples:")
    for r in sorted(b["rows"], key=lambda r: -int(r[h["# Samples"]] or 0))[:top]:
        print("  ", r[h["# Samples"]], r[h["Instructions Executed"]], r[h["Source"]].strip()[:90],
              "| wait", r[h["stall_wait"]], "math", r[h["stall_math"]], "short", r[h["stall_short_sb"]], "long", r[h["stall_long_sb"]], "bar", r[h["stall_barrier"]])
