"""Dynamic (executed) warp-instruction counts per opcode of the first kernel in an ncu report, for the whole kernel and
for each stretch of SASS between barriers, from the report's source page (--import-source on).
python scripts/ncu_opcodes.py file.ncu-rep"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
total = collections.Counter()
segs, cur = [], {"first": 0, "n": 0, "inst": 0, "samp": 0, "ops": collections.Counter()}
for n, r in enumerate(rows[1:]):
    src = r[isrc].strip()
    tok = src.split()
    op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
    try:
        e, s = int(r[iex] or 0), int(r[ismp] or 0)
    except ValueError:
        continue
    total[op] += e
    cur["n"] += 1; cur["inst"] += e; cur["samp"] += s; cur["ops"][op] += e
    if op in ("BAR", "EXIT") or "TRYWAIT" in src:
        cur["end"] = src[:48]
        segs.append(cur)
        cur = {"first": n + 1, "n": 0, "inst": 0, "samp": 0, "ops": collections.Counter()}
segs.append(cur)
tot, tots = sum(total.values()), sum(s["samp"] for s in segs)
print("executed warp-instructions: %d" % tot)
print("-- per opcode, whole kernel")
for k, v in total.most_common(24):
    print("   %-8s %12d  %5.1f %%" % (k, v, 100.0 * v / tot))
fp = sum(total[k] for k in ("FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL"))
print("   fp32 arithmetic (FFMA2 FADD2 FMUL2 FFMA FADD FMUL): %.1f %%" % (100.0 * fp / tot))
print("-- per stretch of SASS between barriers (static line range, share of executed instructions, share of stall samples)")
for s in segs:
    if s["inst"] < 0.003 * tot and s["samp"] < 0.003 * tots:
        continue
    top = ", ".join("%s %.0f%%" % (k, 100.0 * v / max(1, s["inst"])) for k, v in s["ops"].most_common(7))
    print("   #%5d +%4d  inst %5.1f %%  samples %5.1f %%  ends with %-40s | %s" % (s["first"], s["n"], 100.0 * s["inst"] / tot, 100.0 * s["samp"] / max(1, tots), s.get("end", ""), top))
