"""Per-launch DRAM traffic table from an ncu csv (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum).
    python tools/summarize_dram.py in.csv "header comment" > profiles/<name>.csv"""
import csv, io, collections, sys
text = open(sys.argv[1]).read()
rd = csv.DictReader(io.StringIO(text[text.find('"ID"'):]))
per = collections.OrderedDict()
for r in rd:
    d = per.setdefault(r['ID'], {})
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    if r['Metric Name'].startswith('dram'):
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    elif r['Metric Name'].startswith('gpu__time'):
        v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}[u]
    d[r['Metric Name']] = v
n = len(per)
rd_ = sum(d['dram__bytes_read.sum'] for d in per.values())
wr = sum(d['dram__bytes_write.sum'] for d in per.values())
t = sum(d['gpu__time_duration.sum'] for d in per.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print("id,time_us,dram_read_MB,dram_write_MB")
for k, d in per.items():
    print("%s,%.1f,%.2f,%.2f" % (k, d['gpu__time_duration.sum'], d['dram__bytes_read.sum'] / 1e6, d['dram__bytes_write.sum'] / 1e6))
print("TOTAL(%d launches),%.1f,%.1f,%.1f" % (n, t, rd_ / 1e6, wr / 1e6))
