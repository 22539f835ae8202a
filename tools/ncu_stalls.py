"""Per-instruction warp-stall summary of one kernel of an `ncu --set full --import-source on` report.

  ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv
  python tools/ncu_stalls.py /tmp/src.csv "<substring of the kernel name>" [top lines]

Prints the sampled stall reasons summed over the kernel, the samples per opcode, and the hottest SASS lines with their
two main stall reasons (the summaries kept under profiles/*_stalls.txt).
"""
import csv, sys, collections
want = sys.argv[2]
rows = list(csv.reader(open(sys.argv[1])))
# split into kernels
i = 0; blocks = []
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i+1]; j = i+2
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"): j += 1
        blocks.append((name, hdr, rows[i+2:j])); i = j
    else: i += 1
for name, hdr, body in blocks:
    if want not in name: continue
    ix = {h: k for k, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter(); byop = collections.Counter(); nsamp = 0; ninst = 0
    opsamp = collections.Counter()
    for r in body:
        if len(r) < len(hdr): continue
        try: s = int(r[ix["# Samples"]] or 0)
        except: continue
        nsamp += s
        ninst += int(r[ix["Instructions Executed"]] or 0)
        op = r[ix["Source"]].split()[0] if r[ix["Source"]] else "?"
        if op.startswith("@"): op = r[ix["Source"]].split()[1]
        op = op.split(".")[0]
        opsamp[op] += s
        for c in stall_cols:
            v = int(r[ix[c]] or 0); tot[c] += v
    print(name[:60], "samples", nsamp, "inst", ninst)
    print(" stalls:", [(k, v) for k, v in tot.most_common(8)])
    print(" by opcode:", opsamp.most_common(14))
    # top lines
    top = sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0) if len(r) >= len(hdr) and (r[ix["# Samples"]] or "0").isdigit() else 0)[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]
    for r in top:
        st = sorted([(int(r[ix[c]] or 0), c) for c in stall_cols], reverse=True)[:2]
        print("  ", r[ix["Address"]][-5:], r[ix["Source"]][:70], r[ix["# Samples"]], st)
    break
