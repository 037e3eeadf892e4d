#!/usr/bin/env python
"""Per-source-line executed-instruction profile: joins `ncu --page source --csv` (SASS rows, in
address order) with `nvdisasm -gi` line info of the same kernel.
Usage: python profiles/line_profile.py <ncu-rep> <lib.so> <cubin-name-substring> <kernel-substring> [units]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter


def sass_lines(so, cubin_sub, kernel_sub):
    d = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if cubin_sub in f and 'sm_100a' in f][0] if any(cubin_sub in f for f in os.listdir(d)) else sorted(os.listdir(d))[0]
    txt = subprocess.run(['nvdisasm', '-gi', os.path.join(d, cub)], capture_output=True, text=True).stdout.split('\n')
    out, infn, loc = [], False, []
    pending = []
    for ln in txt:
        if ln.startswith('//---') and '.text.' in ln:
            infn = kernel_sub in ln
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            if pending:
                loc = pending
            pending = []
            out.append((int(m.group(1), 16), m.group(2).strip(), list(loc)))
    return out


def main():
    rep, so, cubsub, ksub = sys.argv[1:5]
    units = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ci = {n: i for i, n in enumerate(hdr)}
    prof = [(r[ci['Source']].strip(), int(r[ci['Instructions Executed']] or 0), int(r[ci['# Samples']] or 0))
            for r in rows[2:] if len(r) >= len(hdr)]
    sass = sass_lines(so, cubsub, ksub)
    assert len(sass) == len(prof), (len(sass), len(prof))
    inner, outer, sinner, souter = Counter(), Counter(), Counter(), Counter()
    tot = stot = 0
    for (off, txt, loc), (ptxt, n, smp) in zip(sass, prof):
        tot += n
        stot += smp
        if loc:
            inner['%s:%d' % loc[0]] += n
            outer['%s:%d' % loc[-1]] += n
            sinner['%s:%d' % loc[0]] += smp
            souter['%s:%d' % loc[-1]] += smp
    print('total warp-instructions %d (%.1f per unit); %d stall samples' % (tot, tot / units, stot))
    for name, c, sc in (('innermost source line', inner, sinner), ('outermost (kernel body) line', outer, souter)):
        print('--- by %s: %%instructions, instructions/unit, %%stall samples (~time)' % name)
        keys = sorted(set(k for k, _ in c.most_common(24)) | set(k for k, _ in sc.most_common(24)),
                      key=lambda k: -sc[k])
        for k in keys[:30]:
            print('  %5.1f%%  %9.1f/unit  %5.1f%%  %s' % (100.0 * c[k] / max(tot, 1), c[k] / units,
                                                        100.0 * sc[k] / max(stot, 1), k))


if __name__ == '__main__':
    main()
