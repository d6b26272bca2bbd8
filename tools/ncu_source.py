"""Per-instruction view of an .ncu-rep source page (SASS): opcode mix by executed count and the hottest lines."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
body = [r for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
tot = sum(int(r[iex]) for r in body)
mix = collections.Counter()
for r in body:
    op = r[isrc].split()[0] if not r[isrc].startswith('@') else r[isrc].split()[1]
    mix[op.split('.')[0]] += int(r[iex])
print('total warp-instructions', tot)
for k, v in mix.most_common(18):
    print(f'  {k:10s} {v/tot*100:5.1f}%')
tots = sum(int(r[isamp]) for r in body)
print('hottest by samples:')
for r in sorted(body, key=lambda r: -int(r[isamp]))[:top]:
    print(f'  {int(r[isamp])/tots*100:5.1f}%  ex={int(r[iex]):>10d}  {r[ia][-5:]}  {r[isrc][:90]}')
