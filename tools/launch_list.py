#!/usr/bin/env python
"""Per-kernel table of an `ncu --csv --metrics ...` launch list. Usage: launch_list.py file.csv [first_id]"""
import csv, re, sys, collections
def load(path):
    hdr = None; data = collections.OrderedDict()
    for r in csv.reader(open(path)):
        if r and r[0] == 'ID': hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            key = (int(d['ID']), re.sub(r'^void ', '', re.sub(r'\(.*', '', d['Kernel Name'])).replace('noa_b200::', ''))
            v = float(d['Metric Value'].replace(',', ''))
            u = d['Metric Unit']
            if u in ('us', 'usecond'): v *= 1e3
            if u in ('ms', 'msecond'): v *= 1e6
            if u == 'Kbyte': v *= 1e3
            if u == 'Mbyte': v *= 1e6
            if u == 'Gbyte': v *= 1e9
            data.setdefault(key, {})[d['Metric Name']] = v
    return data
if __name__ == '__main__':
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for (i, name), m in load(sys.argv[1]).items():
        if i < first or name.startswith('at::'): continue
        print(f"{i:4d} {name[:44]:44s} " + ' '.join(f"{k.split('.')[0].replace('sm__inst_executed','inst').replace('gpu__time_duration','ns')}={v:.4g}" for k, v in m.items()))
