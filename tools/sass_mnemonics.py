"""Counts the Blackwell-specific SASS mnemonics per tensor-core kernel of the built library (cuobjdump -sass):
usage: python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "music-fader-nets_b200", "lib", "libfadernets_b200.so")
WANT = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "MUFU.TANH", "MUFU.EX2", "ELECT",
        "MEMBAR", "RED", "REDG", "LDG.E.ENL2.256", "STG.E.ENL2.256")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = {}
cur, counts = None, {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + "."):
                # keep the qualifiers that matter (.2CTA, .MULTICAST, .F16), drop the rest
                key = w + "".join(q for q in (".2CTA", ".MULTICAST", ".F16") if q in op[len(w):])
                counts[cur][key] += 1
                break
dem = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass music-fader-nets_b200/lib/libfadernets_b200.so: Blackwell-specific mnemonics per kernel (tools/sass_mnemonics.py)")
print("# UTCHMMA(.2CTA) = tcgen05.mma (cta_group::2), UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA load (.MULTICAST = cluster multicast),")
print("# UTMASTG = TMA store (cp.async.bulk.tensor shared -> global), SYNCS = mbarrier ops, UCGABAR = cluster barrier, MUFU.TANH(.F16) = tanh.approx.f32 / .f16x2")
tot = collections.Counter()
for mangled, name in sorted(zip(counts, dem), key=lambda kv: kv[1]):
    c = counts[mangled]
    if not any(k.startswith("UTCHMMA") for k in c):
        continue
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    print(name[:160])
    print("    " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
    tot.update(c)
print("# totals over all tensor-core kernels of the library: " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())))
