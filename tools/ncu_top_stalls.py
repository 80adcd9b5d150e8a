#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: top stall sites."""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    si, src, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    data = []
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name':
            break
        if len(r) <= si or r[0] == 'Address':
            continue
        try:
            data.append((int(r[si]), r[src].strip(), int(r[ie])))
        except ValueError:
            pass
    tot = sum(d[0] for d in data) or 1
    print(rows[0][1][:100])
    print('total samples', tot, 'sass instructions', len(data))
    idx = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
    for i in sorted(idx):
        print(f"{i:5d} {data[i][0]:6d} {100 * data[i][0] / tot:5.1f}%  exec={data[i][2]:8d}  {data[i][1][:110]}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
