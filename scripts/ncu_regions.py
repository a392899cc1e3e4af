"""Dev helper: per-code-region stall summary of one kernel from an .ncu-rep (source page, SASS level).
    python scripts/ncu_regions.py gpurun_out/prof.ncu-rep [block_shift=9]"""
import collections, csv, subprocess, sys

path = sys.argv[1]
shift = int(sys.argv[2]) if len(sys.argv) > 2 else 9
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]; data = rows[2:]
ia, isrc, ins, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[ins] or 0) for r in data)
print("kernel:", rows[0][1], " total samples", tot, " SASS instrs", len(data))
blk = collections.OrderedDict()
for r in data:
    a = (int(r[ia], 16) & 0xFFFFFF) >> shift
    d = blk.setdefault(a, [0, 0, collections.Counter(), None])
    d[0] += int(r[ins] or 0); d[1] += int(r[iex] or 0)
    for s in stalls:
        v = int(r[h.index(s)] or 0)
        if v: d[2][s] += v
for a, (n, ex, c, _) in blk.items():
    if n > tot / 400:
        print(f"{a << shift:#x} samples {n:6d} ({100 * n / tot:4.1f}%) instr {ex:10d}", [(k[6:], v) for k, v in c.most_common(4)])
print("top instructions:")
for r in sorted(data, key=lambda r: -int(r[ins] or 0))[:25]:
    print(r[ia][-6:], r[ins], "ex", r[iex], r[isrc][:100])
