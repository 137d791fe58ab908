"""Top warp-stall source lines of an `ncu --set full --import-source on` report: python scripts/ncu_stalls.py rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print(f"# {rep}: {tot} samples, {len(data)} SASS lines")
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]
for i in sorted(top):
    r = data[i]
    st = sorted(((hdr[c], int(r[c])) for c in stall_cols if int(r[c]) > 0), key=lambda kv: -kv[1])[:3]
    print(f"{i:5d} {int(r[isamp]):6d} {int(r[iex]):9d}  {r[ia].strip()[:64]:64s} {st}")
