#!/usr/bin/env python
"""Where do a warp-specialised kernel's cycles go? Reads the SASS source page of an `ncu --set full --import-source on` capture and
prints, in program order, the warp-state samples between the instructions that mark a role's phases (mbarrier waits, TMA / tcgen05
issue points, TMEM loads, bulk stores), plus per-segment totals by opcode and stall reason.

    python tools/ncu_roles.py gpurun_out/prof.ncu-rep [--segments a:b:name,...]
"""
import argparse
import collections
import csv
import io
import subprocess

MARKS = ("SYNCS", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "BAR.", "EXIT", "UTCBAR", "MEMBAR", "FENCE", "DEPBAR", "LDG", "WARPSYNC")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--launch", type=int, default=0, help="which captured launch of the report")
    ap.add_argument("--segments", default="", help="a:b:name,... instruction index ranges to summarise by opcode / stall reason")
    args = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # one block per captured launch: a kernel-name row, the header row, then one row per SASS instruction
    starts = [i for i, r in enumerate(rows) if i + 1 < len(rows) and "Source" in rows[i + 1] and "# Samples" in rows[i + 1]]
    a = starts[args.launch]
    b = starts[args.launch + 1] if args.launch + 1 < len(starts) else len(rows)
    rows = rows[a:b]
    print(rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr, data = rows[1], [r for r in rows[2:] if len(r) == len(rows[1])]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[isamp]) for r in data)
    print(f"total samples {total}, {len(data)} SASS instructions")
    print("index  samples-since-previous-mark  samples-on-it  warp-instr-executed  instruction")
    acc = 0
    for i, r in enumerate(data):
        n = int(r[isamp])
        acc += n
        if any(k in r[isrc] for k in MARKS) and (acc >= total // 400 or "SYNCS.PHASECHK" in r[isrc]):
            print(f"{i:5d} {acc:6d} {n:6d} {r[iex]:>9s}  {r[isrc].strip()[:110]}")
            acc = 0
    for spec in [s for s in args.segments.split(",") if s]:
        a, b, name = spec.split(":")
        a, b = int(a), int(b)
        tot, ex = 0, 0
        byop, bystall = collections.Counter(), collections.Counter()
        for r in data[a:b]:
            n = int(r[isamp])
            tot += n
            ex += int(r[iex])
            tok = r[isrc].split()
            op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
            byop[op] += n
            for c in stall_cols:
                bystall[hdr[c]] += int(r[c])
        print(f"segment {name} [{a},{b}): samples {tot} ({100.0 * tot / total:.1f} %), warp-instructions executed {ex}")
        print("   by opcode:", ", ".join(f"{k} {v}" for k, v in byop.most_common(12)))
        print("   by stall :", ", ".join(f"{k} {v}" for k, v in bystall.most_common(8)))


if __name__ == "__main__":
    main()
