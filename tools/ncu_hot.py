#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv --print-source sass` export, with the dominant stall reason.
    ncu -i rep.ncu-rep --page source --csv --print-source sass > x.csv;  python tools/ncu_hot.py x.csv [top]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    si, src, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for r in rows[2:]:
        if len(r) <= si or r[0] == 'Address':
            continue
        try:
            n = int(r[si])
        except ValueError:
            continue
        stalls = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
        data.append((n, r[src].strip(), int(r[ie] or 0), stalls))
    tot = sum(d[0] for d in data) or 1
    print(rows[0][1][:120])
    print('total samples', tot, 'sass instructions', len(data))
    idx = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
    for i in sorted(idx):
        n, s, ex, st = data[i]
        print(f"{i:5d} {n:7d} {100 * n / tot:5.1f}%  exec={ex:9d}  {s[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
