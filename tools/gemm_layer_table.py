"""ncu launch list restricted to the GEMM kernels of one forward+extract (bench.py --utterances 32
--chunk 192) -> per-layer table: time, DRAM bytes, TFLOP/s, GB/s, and the HBM / tensor floor."""
import collections, csv, re, sys
path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 192
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
recs = collections.OrderedDict()
for row in csv.DictReader(lines):
    d = recs.setdefault(row['ID'], {'by': 0.0})
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    if 'time' in row['Metric Name']:
        d['us'] = v / 1e3 if u == 'ns' else (v if u == 'us' else v * 1e3)
    else:
        d['by'] += v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
seq = []
def blk(tok, C):
    M = n * tok
    return [('qkv', M, 3 * C, C), ('proj', M, C, C), ('lin1', M, 4 * C, C), ('lin2', M, C, 4 * C)]
enc = [(16384, 32, 1), (4096, 64, 2), (1024, 128, 8), (256, 256, 8), (64, 512, 2)]
def encoder(tag):
    for s, (tok, C, nb) in enumerate(enc):
        for b in range(nb):
            for nm, M, N, K in blk(tok, C): seq.append(('%s%d.%s' % (tag, s, nm), M, N, K))
        if s < 4: seq.append(('%s.down%d' % (tag, s), n * tok // 4, 2 * C, 16 * C))
encoder('enc')
for s, (tok, C, nb) in enumerate([(256, 512, 8), (1024, 256, 8), (4096, 128, 2), (16384, 64, 1)]):
    seq.append(('up%d' % s, n * tok // 4, 2 * C, [1024, 512, 256, 128][s]))
    for b in range(nb):
        for nm, M, N, K in blk(tok, C): seq.append(('dec%d.%s' % (s, nm), M, N, K))
encoder('ext')
agg = collections.OrderedDict()
for d, (nm, M, N, K) in zip(recs.values(), seq):
    a = agg.setdefault(nm, [0, 0.0, 0.0, M, N, K]); a[0] += 1; a[1] += d['us']; a[2] += d['by']
tot = floor = 0
print("| layer | cnt | M | N | K | us | DRAM MB | TFLOP/s | GB/s | floor us |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for nm, (c, us, by, M, N, K) in agg.items():
    us /= c; by /= c; fl = 2.0 * M * N * K
    fl_us = max(fl / 1400e6, by / 6.5e6)
    tot += us * c; floor += fl_us * c
    print("| %s | %d | %d | %d | %d | %.1f | %.1f | %.0f | %.0f | %.1f |" % (nm, c, M, N, K, us, by / 1e6, fl / us / 1e6, by / us / 1e3, fl_us))
print("\ntotal %.2f ms, floor (max of 1400 TFLOP/s, 6.5 TB/s on measured DRAM bytes) %.2f ms, %d launches" % (tot / 1e3, floor / 1e3, len(recs)))
