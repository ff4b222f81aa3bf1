"""Top stall-sample SASS lines of one kernel in an .ncu-rep:  python scripts/ncu_top_stalls.py rep kernel_index [n]"""
import csv, subprocess, sys
rep, kid, n = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
seen, data = set(), []
for r in rows[hi + 1:]:
    if len(r) <= i_s or r[0] in seen:
        continue
    seen.add(r[0])
    try:
        data.append((int(r[i_s]), r))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print(rows[0][1][:120] if len(rows[0]) > 1 else "", "| total samples", tot)
for cnt, r in sorted(data, key=lambda x: -x[0])[:n]:
    why = sorted(((int(r[i]), hdr[i][6:]) for i in stall if r[i] not in ("0", "", "-")), reverse=True)[:2]
    print(f"{cnt:6d} {100*cnt/tot:5.1f}% ex={r[i_ex]:>8s} {r[i_src].strip()[:70]:70s} {why}")
