#!/usr/bin/env python3
"""Top stalled SASS instructions of one launch in an ncu report.
usage: ncu_stalls.py <report> <kernel-base-name> <launch-skip> [top N]"""
import csv, io, subprocess, sys
rep, name, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{name}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:120])
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index('# Samples')].isdigit()]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(d[ci['# Samples']]) for d in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(d[ci[h]]) for d in data) for h in stalls}
print('by reason:', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for n, d in sorted(enumerate(data), key=lambda nd: -int(nd[1][ci['# Samples']]))[:top_n]:
    s = {h.replace('stall_', ''): int(d[ci[h]]) for h in stalls if int(d[ci[h]]) > 0}
    s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
    print(str(n).rjust(5), d[ci['# Samples']].rjust(6), d[ci['Source']].strip()[:64].ljust(64), s)
