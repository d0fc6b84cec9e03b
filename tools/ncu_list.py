"""Per-kernel duration (+ optional instruction count) of the LAST iteration in an `ncu --csv` launch list."""
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
anchor = sys.argv[2] if len(sys.argv) > 2 else 'encode_hist'
byid = {}
for x in csv.DictReader(lines):
    byid.setdefault(int(x['ID']), {'name': x['Kernel Name'][:64], 'grid': x['Grid Size']})[x['Metric Name']] = float(x['Metric Value'].replace(',', ''))
ids = sorted(byid)
last = [i for i in ids if anchor in byid[i]['name']][-1]
for i in ids:
    k = byid[i]
    if i >= last - 1 and 'ss::' in k['name']:
        print(f"{k['gpu__time_duration.sum'] / 1000:7.1f} us {k.get('smsp__inst_executed.sum', 0) / 1e6:7.2f} Minst {k['grid']:>14}  {k['name']}")
