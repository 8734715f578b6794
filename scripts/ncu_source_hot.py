"""Per-CUDA-source-line instruction / stall-sample shares from `ncu -i rep --page source --csv --print-source cuda,sass`.
usage: ncu_source_hot.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
funcs, cur = [], None
for r in rows:
    if r and r[0] == "Function Name":
        cur = {"name": r[1], "lines": []}
        funcs.append(cur)
    elif cur is not None and r and r[0].isdigit() and len(r) > 8 and r[2] == "-":
        cur["lines"].append((int(r[0]), r[1], int(r[7] or 0), int(r[6] or 0)))
for f in funcs:
    tot, ts = sum(l[2] for l in f["lines"]), sum(l[3] for l in f["lines"])
    if tot < 1e6:
        continue
    print(f["name"], "instr", tot, "samples", ts)
    for ln, src, ins, sm in sorted(f["lines"], key=lambda l: -l[2])[:top]:
        print(f"  {ln:5d} {100 * ins / tot:5.1f}% ins {100 * sm / max(ts, 1):5.1f}% smp  {src[:120]}")
