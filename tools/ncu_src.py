"""Summarise an ncu report's source page: stall reasons and the hottest SASS lines. usage: ncu_src.py rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']  # one source table per profiled kernel: take the longest-running (last) one
hi = his[-1]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}; data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != 'Address']
tot = sum(int(r[ix['# Samples']]) for r in data)
print('samples', tot, 'warp instr', sum(int(r[ix['Instructions Executed']]) for r in data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print(sorted(agg.items(), key=lambda x: -x[1])[:9])
for n, r in enumerate(data): r.append(n)
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:topn]:
    st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(str(r[-1]).rjust(5), r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), r[ix['Source']].strip()[:64].ljust(64), st)
