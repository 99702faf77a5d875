#!/usr/bin/env python
"""Per-region view of an ncu source page: python profiles/source_regions.py <report.ncu-rep> [chunk]
Groups the kernel's SASS into chunks of `chunk` instructions and prints each chunk's share of executed instructions and of
warp samples, the marker instructions it contains and its top stall reasons -- enough to tell the roles of a warp-specialised
kernel apart (producer / MMA issuer / epilogue / finisher) without opening the GUI."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 10]
    iS, iN, iI = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot_i = sum(int(r[iI]) for r in data) or 1
    tot_s = sum(int(r[iN]) for r in data) or 1
    print("kernel: %s   instructions %d   samples %d" % (rows[0][1] if len(rows[0]) > 1 else "?", tot_i, tot_s))
    keys = ("UTCHMMA", "UTCQMMA", "UBLKCP", "UTMALDG", "LDTM", "STTM", "MUFU.LG2", "MUFU.EX2", "BAR.SYNC", "UTCBAR", "SHFL", "TRYWAIT", "LDG", "STG", "LDS", "STS", "ATOM", "RED")
    for a in range(0, len(data), chunk):
        seg = data[a:a + chunk]
        ins = sum(int(r[iI]) for r in seg)
        sm = sum(int(r[iN]) for r in seg)
        if ins == 0 and sm == 0:
            continue
        found = []
        for k in keys:
            n = sum(1 for r in seg if k in r[iS])
            if n:
                found.append("%s x%d" % (k, n))
        st = {}
        for r in seg:
            for i in stall_cols:
                if r[i] and r[i] != "0":
                    st[h[i]] = st.get(h[i], 0) + int(r[i])
        top = sorted(st.items(), key=lambda x: -x[1])[:3]
        print("%5d-%5d  inst %5.1f%%  samples %5.1f%%  | %s | %s" % (a, a + len(seg), 100.0 * ins / tot_i, 100.0 * sm / tot_s,
                                                                  " ".join("%s=%.1f%%" % (k[6:], 100.0 * v / tot_s) for k, v in top), ", ".join(found)))


if __name__ == "__main__":
    main()
