#!/usr/bin/env python
"""Development aid: where a kernel's executed instructions go, from the SASS source page of an ncu --set full --import-source on
capture:  python tools/ncu_source_budget.py gpurun_out/x.ncu-rep <units> > profiles/rN_x_instruction_budget.txt
`units` = the number of work units of the launch (dispatch records for the meshlet test kernel); the report is per unit.
Prints the opcode histogram (executed warp instructions and stall samples) and the same over consecutive blocks of the
SASS listing, for the LAST launch in the report."""
import collections
import csv
import io
import subprocess
import sys


def main():
    path, units = sys.argv[1], float(sys.argv[2])
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    kernels, cur = [], None
    for r in csv.reader(io.StringIO(raw)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif cur is not None and r:
            cur["rows"].append(r)
    k = kernels[-1]
    hdr, body = k["rows"][0], k["rows"][1:]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot = sum(int(r[iex]) for r in body)
    tsm = sum(int(r[ismp]) for r in body) or 1
    print("# %s" % k["name"])
    print("# %d SASS instructions, %d executed warp instructions = %.1f per unit (%d units)" % (len(body), tot, tot / units, units))

    def op_of(r):
        s = r[isrc].strip()
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        return s.split()[0]
    h, hs = collections.Counter(), collections.Counter()
    for r in body:
        op = op_of(r).split(".")[0]
        h[op] += int(r[iex]); hs[op] += int(r[ismp])
    print("# opcode      executed   per unit   share  stall samples")
    for op, c in h.most_common(40):
        print("%-10s %10d  %8.1f  %5.1f%%  %5.1f%%" % (op, c, c / units, 100.0 * c / tot, 100.0 * hs[op] / tsm))
    print("# blocks of 48 SASS instructions: first index, share of stall samples, share of executed instructions (per unit), most frequent opcodes")
    for i in range(0, len(body), 48):
        blk = body[i:i + 48]
        e = sum(int(r[iex]) for r in blk)
        if e == 0:
            continue
        s = sum(int(r[ismp]) for r in blk)
        ops = collections.Counter(op_of(r) for r in blk)
        print("%5d  %5.1f%%  %5.1f%% (%6.1f)  %s" % (i, 100.0 * s / tsm, 100.0 * e / tot, e / units, " ".join("%s:%d" % oc for oc in ops.most_common(6))))


if __name__ == "__main__":
    main()
