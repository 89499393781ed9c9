#!/usr/bin/env python
"""hot SASS instructions of an ncu report (source page): python tools/ncu_hot.py rep.ncu-rep [kernel regex] [top N]"""
import csv, re, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else "."; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        i = j
        if not re.search(pat, name):
            continue
        print("==", name[:140])
        c = {h: k for k, h in enumerate(hdr)}
        S, X, W, WI = c["# Samples"], c["Instructions Executed"], c["L1 Wavefronts Shared"], c["L1 Wavefronts Shared Ideal"]
        tot = sum(int(r[S]) for r in body); totx = sum(int(r[X]) for r in body); totw = sum(int(r[W]) for r in body)
        print("   samples %d, warp instructions %d, shared wavefronts %d (ideal %d)" % (tot, totx, totw, sum(int(r[WI]) for r in body)))
        mix = {}
        for r in body:
            op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
            op = op.split(".")[0]
            mix[op] = mix.get(op, 0) + int(r[X])
        print("   mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / totx) for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:18]))
        print("   %-8s %-8s %-8s %-9s %s" % ("samp%", "exec%", "wavef%", "wf/ideal", "sass"))
        for r in sorted(body, key=lambda r: -int(r[S]))[:top]:
            w, wi = int(r[W]), int(r[WI])
            print("   %-8.2f %-8.2f %-8.2f %-9s %s" % (100.0 * int(r[S]) / tot, 100.0 * int(r[X]) / totx, 100.0 * w / max(1, totw), ("%.2f" % (w / wi)) if wi else "-", r[1].strip()[:90]))
    else:
        i += 1
