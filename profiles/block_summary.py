"""Aggregate an `ncu --page source --csv` export by basic block (hot loops vs overhead)."""
import csv
import re
import sys


def main(path, top=12):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    uniq, seen = [], set()
    for r in rows[2:]:
        if len(r) < len(hdr) or r[ix['Address']] in seen:
            continue
        seen.add(r[ix['Address']])
        uniq.append(r)

    def f(r, k):
        try:
            return float(r[ix[k]])
        except ValueError:
            return 0.0
    tot = sum(f(r, '# Samples') for r in uniq)
    te = sum(f(r, 'Instructions Executed') for r in uniq)
    blocks, cur = [], []
    for r in uniq:
        cur.append(r)
        op = re.sub(r'^@!?U?P\d+\s+', '', r[ix['Source']].strip()).split()[0]
        if op.startswith(('BRA', 'BSYNC', 'CALL', 'RET', 'EXIT', 'BSSY', 'WARPSYNC')):
            blocks.append(cur)
            cur = []
    blocks.append(cur)
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    print('total samples', tot, 'warp instructions', te)
    for b in sorted(blocks, key=lambda b: -sum(f(r, '# Samples') for r in b))[:top]:
        sm = sum(f(r, '# Samples') for r in b)
        ex = sum(f(r, 'Instructions Executed') for r in b)
        nd = sum('DMMA' in r[ix['Source']] for r in b)
        st = {k[6:]: round(100 * sum(f(r, k) for r in b) / max(sm, 1)) for k in stalls}
        st = {k: v for k, v in st.items() if v >= 3}
        print('%s len=%d dmma=%d samples=%.1f%% exec=%.1f%% first_exec=%d %s' % (
            b[0][ix['Address']][-5:], len(b), nd, 100 * sm / tot, 100 * ex / te, f(b[0], 'Instructions Executed'), st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12)
