"""Aggregate an `ncu --page source --csv` export: stall reasons, opcodes, and hottest code regions.
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME --launch-skip i --launch-count 1 > src.csv
    python profiles/stall_summary.py src.csv
"""
import collections
import csv
import re
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen, uniq = set(), []
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[ix['Address']].startswith('0x') and not re.match(r'^[0-9a-f]+$', r[ix['Address']]):
            continue
        if r[ix['Address']] in seen:
            continue
        seen.add(r[ix['Address']])
        uniq.append(r)

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0
    tot = sum(f(r, '# Samples') for r in uniq)
    te = sum(f(r, 'Instructions Executed') for r in uniq)
    print("kernel:", rows[0][1][:90])
    print("instructions %d, samples %d, warp instructions executed %.4g" % (len(uniq), tot, te))
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {k: sum(f(r, k) for r in uniq) for k in stalls}
    print("stall reasons (% of samples):", {k[6:]: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v / tot > 0.004})
    byop, exe = collections.Counter(), collections.Counter()
    for r in uniq:
        op = re.sub(r'^@!?U?P\d+\s+', '', r[ix['Source']].strip()).split()[0].split('.')[0]
        byop[op] += f(r, '# Samples')
        exe[op] += f(r, 'Instructions Executed')
    print("samples by opcode (%):", [(k, round(100 * v / tot, 1)) for k, v in byop.most_common(14)])
    print("executed by opcode (%):", [(k, round(100 * v / te, 1)) for k, v in exe.most_common(16)])
    print("hottest instructions:")
    for r in sorted(uniq, key=lambda r: -f(r, '# Samples'))[:top]:
        s = ' '.join('%s=%d' % (k[6:], f(r, k)) for k in stalls if f(r, k) > 0.1 * f(r, '# Samples'))
        print('  %5.2f%% %s exec=%-9d %-52s %s' % (100 * f(r, '# Samples') / tot, r[ix['Address']][-5:],
                                                 f(r, 'Instructions Executed'), r[ix['Source']].strip()[:52], s))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
