"""Stall samples of the warp-specialised GEMM split by warp role (address ranges between the role's marker instructions):
python scripts/ncu_role_samples.py rep kernel_index"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]; s = h.index("Source"); n = h.index("# Samples"); e = h.index("Instructions Executed")
stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
seen, ins = set(), []
for r in rows[hi + 1:]:
    if r[0] in seen or not r[0].startswith("0x"):
        continue
    seen.add(r[0])
    ins.append(r)
# role boundaries: first UTMALDG.*2CTA-or-3D (producer), first epilogue-side tmem wait, first UTCHMMA (MMA)
def first(pred, start=0):
    return next(i for i in range(start, len(ins)) if pred(ins[i][s]))
p0 = first(lambda t: "UTMALDG" in t) - 40
m0 = first(lambda t: "UTCHMMA" in t) - 60
e0 = first(lambda t: "LDTM" in t) - 80
bounds = sorted([(0, "prologue"), (p0, "producer"), (e0, "epilogue"), (m0, "mma")])
tot = sum(int(r[n]) for r in ins)
for k, (b, name) in enumerate(bounds):
    hi_ = bounds[k + 1][0] if k + 1 < len(bounds) else len(ins)
    seg = ins[b:hi_]
    c = sum(int(r[n]) for r in seg)
    why = {}
    for r in seg:
        for i in stall:
            if r[i] not in ("0", "", "-"):
                why[h[i][6:]] = why.get(h[i][6:], 0) + int(r[i])
    top = sorted(why.items(), key=lambda x: -x[1])[:6]
    print(f"{name:9s} {c:6d} {100*c/tot:5.1f}%  {top}")
