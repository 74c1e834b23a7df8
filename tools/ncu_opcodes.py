"""Executed-instruction histogram by SASS opcode for one kernel launch of an ncu report.
   python tools/ncu_opcodes.py <ncu-rep> <kernel-regex> <launch-skip>"""
import collections, csv, subprocess, sys
rep, kre, skip = sys.argv[1], sys.argv[2], sys.argv[3]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; data = rows[2:]
iS, iE, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
by = collections.Counter(); byT = collections.Counter(); tot = 0
for r in data:
    try:
        e, t = int(r[iE]), int(r[iT])
    except (ValueError, IndexError):
        continue
    toks = r[iS].strip().split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    by[op] += e; byT[op] += t; tot += e
print("total", tot)
for k, v in by.most_common(40):
    print("%6.2f%% %12d thr/inst %5.1f %s" % (100.0 * v / tot, v, byT[k] / max(v, 1), k))
