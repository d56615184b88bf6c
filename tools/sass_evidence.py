"""SASS opcode histogram of the Blackwell-native instructions in libmvg_b200.so -> profiles/sass_evidence_<tag>.txt
    python tools/sass_evidence.py r2          (CPU only: cuobjdump -sass)"""
import collections, os, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "mvgformer_b200", "libmvg_b200.so")],
                     capture_output=True, text=True).stdout
pat = re.compile(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
KEYS = ("UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTCATOMSWS", "HFMA2",
        "ACQBULK", "PREEXIT", "UTMACCTL", "LDS.128")
tot, per, cur = collections.Counter(), collections.OrderedDict(), None
for line in txt.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("mvg::", "")
        continue
    m = pat.match(line)
    if m and cur:
        op = m.group(1).rstrip(";")
        if op.startswith(KEYS):
            if not op.startswith("LDS"):
                tot[op] += 1
            short = op.split(".TRANS")[0] if op.startswith("SYNCS") else op.rstrip(".")
            per.setdefault(cur, collections.Counter())[short] += 1
with open(os.path.join(ROOT, "profiles", f"sass_evidence_{tag}.txt"), "w") as f:
    f.write("cuobjdump -sass mvgformer_b200/libmvg_b200.so | opcode histogram of Blackwell-native instructions (final tree; tools/sass_evidence.py)\n"
            "tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, TMA tensor load/store -> UTMALDG/UTMASTG (.2D and .3D;\n"
            "UTMALDG.3D = the NCHW pyramid levels loaded as an MN-major operand), cp.async.bulk (tile rows + record blocks of the gather) -> UBLKCP,\n"
            "mbarrier -> SYNCS, packed fp16 FMA of the gather blend -> HFMA2, griddepcontrol.wait / launch_dependents (programmatic dependent\n"
            "launch) -> ACQBULK / PREEXIT, prefetch.tensormap -> UTMACCTL.PF\n")
    for op, c in tot.most_common():
        f.write(f"{c:7d} {op}\n")
    f.write("\nper kernel (template instances summed):\n")
    for k, c in per.items():
        if any(x.startswith(("UTC", "UBLKCP", "LDTM", "UTMA", "HFMA2", "ACQBULK")) for x in c):
            f.write(f"  {k}:\n     " + "".join(f"{v:5d} {o};  " for o, v in sorted(c.items())) + "\n")
print(open(os.path.join(ROOT, "profiles", f"sass_evidence_{tag}.txt")).read()[:3000])
