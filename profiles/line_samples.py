"""Warp-stall samples of an `ncu --page source --csv --print-source sass` export aggregated by SOURCE LINE.
The line of every SASS instruction comes from `nvdisasm --print-line-info` on the cubin of the same build
(the .so travels to the GPU box, so offsets agree); ncu addresses are matched by their offset from the first one.

    cuobjdump -xelf all blueice_b200/build/<tu>.o          # -> <tu>.sm_100a.cubin
    nvdisasm --print-line-info <tu>.sm_100a.cubin > dis.txt
    ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
    python profiles/line_samples.py dis.txt '<mangled kernel name substring>' sass.csv [top]
"""
import collections
import csv
import re
import sys


def line_table(dis_path, kernel):
    table, cur, inside = {}, None, False
    for ln in open(dis_path, errors='replace'):
        if ln.startswith('.text.'):
            inside = kernel in ln
            continue
        if ln.startswith('//---------------------') and '.text.' not in ln:
            inside = False
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main(dis_path, kernel, sass_csv, top=40):
    table = line_table(dis_path, kernel)
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    base, seen = None, set()
    by_line = collections.defaultdict(lambda: collections.Counter())
    tot = 0.0
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[ix['Address']].startswith('0x'):
            continue
        a = int(r[ix['Address']], 16)
        if a in seen:
            continue
        seen.add(a)
        if base is None:
            base = a
        loc = table.get(a - base, (None, '?'))[0] or ('?', 0)
        n = float(r[ix['# Samples']] or 0)
        tot += n
        c = by_line[loc]
        c['samples'] += n
        c['exec'] += float(r[ix['Instructions Executed']] or 0)
        c['n_inst'] += 1
        for s in stalls:
            c[s[6:]] += float(r[ix[s]] or 0)
    print("total samples %d over %d source lines (%d instructions mapped of %d)" % (tot, len(by_line), len(seen), len(table)))
    for loc, c in sorted(by_line.items(), key=lambda kv: -kv[1]['samples'])[:top]:
        why = ' '.join('%s=%d%%' % (k, 100 * v / c['samples']) for k, v in c.most_common() if k not in ('samples', 'exec', 'n_inst')
                       and c['samples'] and v > 0.12 * c['samples'])
        print("  %5.2f%%  %-28s inst=%-4d exec=%-10d %s" % (100 * c['samples'] / tot, '%s:%d' % loc, c['n_inst'], c['exec'], why))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
