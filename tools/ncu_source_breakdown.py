"""Per-region stall breakdown of an `ncu --set full --import-source on` capture of shade_tc_kernel (source page):

    python tools/ncu_source_breakdown.py gpurun_out/r2o_render.ncu-rep > profiles/r2o_stall_breakdown.txt

Regions: the mbarrier spin loops (a warp waiting for its partner role), the rolled section loops of the epilogue variants
(found as backward branches whose body contains a tcgen05.ld), everything else.  For every region: share of all warp
samples, instructions executed, and the distribution of stall reasons -- the evidence behind DESIGN.md 4.1 "where the
time goes"."""
import collections
import csv
import re
import subprocess
import sys


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0

    stalls = [k for k in hdr if k.startswith("stall_") and "(" not in k]
    addr = [int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]]) for r in data]
    base, a2i = addr[0], {a: i for i, a in enumerate(addr)}
    src = [r[ix["Source"]].strip() for r in data]
    total = sum(f(r, "# Samples") for r in data)
    region = ["other"] * len(data)
    # spin loops: a SYNCS.PHASECHK try-wait and the few instructions of its retry loop
    for i, s in enumerate(src):
        if "SYNCS.PHASECHK" in s:
            for j in range(max(0, i - 6), min(len(data), i + 6)):
                if f(data[j], "Instructions Executed") > 4 * f(data[max(0, i - 12)], "Instructions Executed") or j >= i:
                    region[j] = "spin: waiting on an mbarrier"
    for i, s in enumerate(src):
        if "EXIT" in s or (i + 1 < len(src) and "EXIT" in src[i + 1]):
            region[i] = "idle warps at the final barrier"
    loops = []
    for i, s in enumerate(src):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t < addr[i] and t in a2i and 60 < i - a2i[t] < 1500 and any("LDTM" in x for x in src[a2i[t]:i + 1]):
                loops.append((a2i[t], i))
    for s0, e0 in loops:
        body = src[s0:e0 + 1]
        mufu = sum("MUFU" in x for x in body)
        kind = ("softplus + softplus' saved" if mufu >= 48 else "softplus") if mufu >= 32 else \
               ("gradient chain" if sum("PRMT" in x for x in body) >= 16 else ("ReLU + rank update / rows" if sum("SHFL" in x for x in body) >= 32 else "ReLU"))
        name = f"section loop @{addr[s0] - base:#x} ({e0 - s0 + 1} instr, {kind})"
        for j in range(s0, e0 + 1):
            if region[j] == "other":
                region[j] = name
    agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    for i, r in enumerate(data):
        a = agg[region[i]]
        a[0] += f(r, "# Samples")
        a[1] += f(r, "Instructions Executed")
        for k in stalls:
            a[2][k] += f(r, k)
    print(f"report: {rep}\ntotal warp samples: {total:.0f}\n")
    print(f"{'region':<78}{'samples':>9}{'share':>8}{'warp instr':>12}  stall reasons (share of the region's samples)")
    for name, (n, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if n < 50:
            continue
        top = "  ".join(f"{k[6:]} {100 * v / max(n, 1):.0f}%" for k, v in st.most_common(6))
        print(f"{name:<78}{n:>9.0f}{100 * n / total:>7.1f}%{ex:>12.3g}  {top}")
    hot = sum(v[0] for k, v in agg.items() if k.startswith("section loop"))
    print(f"\nsection loops together: {100 * hot / total:.1f} % of all samples; 'selected' = the warp issued, 'not_selected' = it was "
          "ready but another warp of the sub-partition issued; long_sb = waiting for a global / local / TMEM load, short_sb = for MUFU / "
          "shared-memory / shuffle results, wait = fixed-latency dependency, mio = MIO queue full, no_inst = instruction fetch.")


if __name__ == "__main__":
    main(sys.argv[1])
