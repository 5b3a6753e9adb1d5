"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
by_samples = len(sys.argv) > 3 and sys.argv[3] == "samples"
hdr = None
agg = collections.defaultdict(lambda: [0, 0, ''])
cur = None
fname = ''
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r; iexec = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); continue
    if hdr is None or len(r) <= iexec: continue
    if r[0] != '':
        cur = (fname, int(r[0])); agg[cur][2] = r[1]; continue
    try:
        ie = int(r[iexec]); sm = int(r[isamp])
    except ValueError:
        continue
    agg[cur][0] += ie; agg[cur][1] += sm
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total warp-inst", tot, "samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_samples else 0])[:top]:
    print("%-14s %4d inst=%5.1f%% samp=%5.1f%%  %s" % (k[0][:14], k[1], 100 * v[0] / tot, 100 * v[1] / max(1, tots), v[2].strip()[:100]))
