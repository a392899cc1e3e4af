"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, from `cuobjdump -sass` of the built library.
    python scripts/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "video_fingerprint_b200", "libvfp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
pats = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                                ("UTMASTG", r"\bUTMASTG"), ("UTMAPF", r"\bUTMAPF"), ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"), ("LDGSTS", r"\bLDGSTS"),
                                ("LDSM", r"\bLDSM"), ("HMMA", r"\bHMMA"), ("MUFU", r"\bMUFU")])
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
print("# cuobjdump -sass video_fingerprint_b200/libvfp_b200.so : instruction counts per kernel (sm_100a)")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = bulk copy,")
print("# UTCBAR = tcgen05.commit, LDGSTS = cp.async, LDSM = ldmatrix, HMMA = mma.sync (legacy tensor path)")
print("%-96s " % "kernel" + " ".join("%8s" % k[:8] for k in pats))
for fn, c in counts.items():
    if not any(c.values()):
        continue
    name = re.sub(r"\(.*", "", demangle(fn)).replace("vfp::", "").replace("void ", "")
    print("%-96s " % name[:96] + " ".join("%8d" % c[k] for k in pats))
