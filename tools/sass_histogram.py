"""Per-kernel SASS evidence of the Blackwell-native paths: counts of UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG / UTMAREDG
(TMA tensor loads / stores / reduce-adds), UBLKCP (cp.async.bulk), LDTM / STTM (tcgen05.ld / st) and legacy HMMA in the
shipped library.  python tools/sass_histogram.py > profiles/r02_sass_histogram.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "pitchextractor_b200", "libpe_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "HMMA", "SYNCS", "ELECT"]
kern, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("pe::", "")
        rows[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1).split(".")[0]
        rows[kern]["_total"] += 1
        if op in KEYS:
            rows[kern][op] += 1
print("# SASS histogram of libpe_b200.so (sm_100a), round 2\n")
print("`cuobjdump -sass pitchextractor_b200/libpe_b200.so`, instruction counts per kernel (static).  UTC*MMA = tcgen05.mma,")
print("UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st,")
print("SYNCS = mbarrier ops, HMMA = legacy mma.sync (none expected).\n")
print("| kernel | instr | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
tot = collections.Counter()
for k, c in rows.items():
    if not any(c[x] for x in KEYS if x not in ("SYNCS", "ELECT")):
        continue
    print("| `%s` | %d | " % (k[:70], c["_total"]) + " | ".join(str(c[x]) if c[x] else "" for x in KEYS) + " |")
    tot.update(c)
print("| **all kernels listed** | %d | " % tot["_total"] + " | ".join(str(tot[x]) for x in KEYS) + " |")
print("\nKernels without any tensor-core / TMA / TMEM instruction (elementwise, reductions, optimizer): %d of %d."
      % (sum(1 for c in rows.values() if not any(c[x] for x in KEYS if x not in ("SYNCS", "ELECT"))), len(rows)))
