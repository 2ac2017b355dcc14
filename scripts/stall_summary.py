#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall samples by reason and the top stalled SASS instructions.
    python scripts/stall_summary.py gpurun_out/prof_X_source.csv [N] [out.md]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    kernel = rows[0][1] if len(rows[0]) > 1 else ''
    hdr = rows[1]
    si, src, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    body = [r for r in rows[2:] if len(r) > si]
    tot = sum(float(r[si]) for r in body) or 1.0
    by_reason = {h: sum(float(r[hdr.index(h)]) for r in body) for h in reasons}
    out = [f'kernel: `{kernel[:100]}`', f'{int(tot)} stall samples over {len(body)} SASS instructions, '
           f'{sum(float(r[ie]) for r in body):.0f} warp instructions executed', '', '| stall reason | share |', '|---|---|']
    rs = sum(by_reason.values()) or 1.0
    for h, v in sorted(by_reason.items(), key=lambda kv: -kv[1])[:8]:
        out.append(f'| {h} | {100 * v / rs:.1f} % |')
    out += ['', '| samples | share | executed | dominant stall | instruction |', '|---|---|---|---|---|']
    for r in sorted(body, key=lambda r: -float(r[si]))[:topn]:
        st = {h: float(r[hdr.index(h)]) for h in reasons}
        out.append(f'| {int(float(r[si]))} | {100 * float(r[si]) / tot:.1f} % | {r[ie]} | {max(st, key=st.get)} | `{r[src].strip()[:80]}` |')
    text = '\n'.join(out)
    print(text)
    if len(sys.argv) > 3:
        open(sys.argv[3], 'w').write(text + '\n')


if __name__ == '__main__':
    main()
