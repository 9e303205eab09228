"""Count the Blackwell-native SASS mnemonics in every object of libst_b200.so's build (UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG = TMA loads, HMMA would be the legacy mma.sync path).  python tools/sass_summary.py > profiles/...txt"""
import collections, glob, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = sorted(glob.glob(os.path.join(root, "speech-tranformer-pytorch_b200", "build", "*.o")))
pats = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "HGMMA", "SYNCS", "ELECT"]
print(f"{'object':22s}" + "".join(f"{p:>9s}" for p in pats) + "   kernels")
for o in objs:
    sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    c = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for p in pats:
                if op.startswith(p): c[p] += 1
    nk = sass.count("Function :")
    print(f"{os.path.basename(o):22s}" + "".join(f"{c[p]:9d}" for p in pats) + f"   {nk}")
