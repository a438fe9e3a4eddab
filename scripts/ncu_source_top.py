"""Summarise an `ncu --page source --csv` export: top SASS lines by stall samples, stall-reason totals."""
import csv, sys
csv.field_size_limit(10**9)
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == 'Address' and 'Source' in r]
# the csv holds (SASS view, source view) per launch; take SASS tables (those whose Source column looks like SASS)
tables = []
for k, s in enumerate(starts):
    e = starts[k + 1] if k + 1 < len(starts) else len(rows)
    tables.append((rows[s], [r for r in rows[s + 1:e] if len(r) == len(rows[s])]))
hdr, body = tables[which]
ix = {n: i for i, n in enumerate(hdr)}
tot = sum(float(r[ix['# Samples']] or 0) for r in body)
inst = sum(float(r[ix['Instructions Executed']] or 0) for r in body)
print(f"table {which}/{len(tables)}: {len(body)} lines, samples={tot:.0f}, warp-instructions executed={inst:.0f}")
stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
st = {n: sum(float(r[ix[n]] or 0) for r in body) for n in stall_cols}
print("stall totals:", ", ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > 0))
body2 = sorted(body, key=lambda r: -float(r[ix['# Samples']] or 0))[:topn]
for r in body2:
    reasons = sorted(((float(r[ix[n]] or 0), n[6:]) for n in stall_cols), reverse=True)[:2]
    print(f"{float(r[ix['# Samples']]):7.0f} {100*float(r[ix['# Samples']])/max(tot,1):5.1f}%  ex={float(r[ix['Instructions Executed']] or 0):9.0f}  {r[ix['Source']][:70]:70s} {reasons[0][1]}:{reasons[0][0]:.0f} {reasons[1][1]}:{reasons[1][0]:.0f}")
# opcode histogram by executed count
from collections import Counter
c = Counter()
for r in body:
    src = r[ix['Source']].strip()
    op = src.split()[0] if src and not src.startswith('@') else (src.split()[1] if len(src.split()) > 1 else src)
    c[op.split('.')[0]] += float(r[ix['Instructions Executed']] or 0)
print("executed by opcode:", ", ".join(f"{k}={v:.0f}" for k, v in c.most_common(18)))
