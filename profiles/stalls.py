"""Top stall sites of a kernel from an ncu report's source page (needs -lineinfo + --import-source on).
usage: python profiles/stalls.py <report.ncu-rep> [top_n]   -> SASS instructions ranked by stall samples, with the
dominant stall reason each; the mbarrier try_wait loops identify which pipeline role waits for which."""
import csv
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    return rows[0][1] if rows[0] else "", hdr, rows[hdr_i + 1:]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    name, hdr, rows = load(rep)
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[col["# Samples"]] or 0) for r in rows)
    print(f"kernel: {name[:110]}\ntotal samples: {tot}")
    ranked = sorted(rows, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]
    for r in ranked:
        n = int(r[col["# Samples"]] or 0)
        reasons = sorted(((int(r[col[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:2]
        rs = ", ".join(f"{k} {v}" for v, k in reasons if v)
        print(f"{100.0 * n / max(tot, 1):5.1f}%  {r[col['Address']][-5:]}  {r[col['Source']][:90]:90s} [{rs}]")


if __name__ == "__main__":
    main()
