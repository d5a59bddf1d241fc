"""Per-opcode and per-region instruction / shared-wavefront breakdown of one kernel from an .ncu-rep
(source page).  Usage: python profiles/ncu_regions.py rep.ncu-rep <cta_steps>   (cta_steps = CTAs x time steps)"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep, steps = sys.argv[1], float(sys.argv[2])
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    iw, iwi, ismp = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(float(r[iex] or 0) for r in data)
    print(f"total warp-instr {tot:.3e} = {tot / steps:.0f} per CTA-step; SASS lines {len(data)}")
    cls, wf, wfi = defaultdict(float), defaultdict(float), defaultdict(float)
    stalls = defaultdict(float)
    for r in data:
        parts = r[isrc].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        if op in ("LDS", "STS", "LDG", "STG"):
            op = (parts[1] if parts[0].startswith("@") else parts[0])
            op = ".".join(op.split(".")[:3]) if op.startswith("LDG") else op
        e = float(r[iex] or 0)
        cls[op] += e; wf[op] += float(r[iw] or 0); wfi[op] += float(r[iwi] or 0)
        for i in stall_cols:
            stalls[hdr[i]] += float(r[i] or 0)
    for c, e in sorted(cls.items(), key=lambda x: -x[1])[:32]:
        print(f"  {c:22s} {e / steps:9.0f} /CTA-step   smem wavefronts {wf[c] / steps:8.0f} (ideal {wfi[c] / steps:8.0f})")
    ts = sum(stalls.values())
    print("stall samples:", ", ".join(f"{k[6:]}={100 * v / ts:.1f}%" for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:9]))
    reg, cur = [], None
    for r in data:
        e = float(r[iex] or 0)
        if cur is None or abs(e - cur[2]) > 0.02 * max(e, cur[2], 1):
            cur = [r[ia], r[ia], e, 0, 0.0]
            reg.append(cur)
        cur[1] = r[ia]; cur[3] += 1; cur[4] += float(r[ismp] or 0)
    for a, b, e, n, s in reg:
        if e * n / steps > 400:
            print(f"  region {a[-5:]}-{b[-5:]}: {n:4d} sass x {e / steps:7.1f} exec/CTA-step = {e * n / steps:8.0f}   samples {s:.0f}")


if __name__ == "__main__":
    main()
