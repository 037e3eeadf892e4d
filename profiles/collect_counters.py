#!/usr/bin/env python
"""Per-workload hardware counters of ONE bench step (all of this repo's kernels in it): executed warp-instructions,
kernel time and DRAM bytes.  Two modes:

  python profiles/collect_counters.py run   [workloads...]   # on the GPU box: ncu around `bench.py --profile-one`
  python profiles/collect_counters.py parse [workloads...]   # anywhere: gpurun_out/cnt_*.csv -> profiles/issue.json,
                                                             #   profiles/traffic.json, profiles/r2_counters.txt
bench.py reads issue.json (roofline.issue) and traffic.json (roofline.traffic).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
ALL = ['dynaq', 'dynaq64k', 'pma', 'q', 'sr', 'sfma', 'sr100']
METRICS = 'smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum'


def run(wls):
    os.makedirs(OUT, exist_ok=True)
    for w in wls:
        cmd = ['ncu', '--profile-from-start', 'off', '--metrics', METRICS, '--clock-control', 'none', '--csv',
               '--log-file', os.path.join(OUT, 'cnt_%s.csv' % w), sys.executable, os.path.join(ROOT, 'bench.py'),
               '--workload', w, '--no-pma', '--no-cpu', '--profile-one']
        with open(os.path.join(OUT, 'cnt_%s.json' % w), 'w') as f:
            subprocess.run(cmd, stdout=f, stderr=subprocess.DEVNULL, check=False)


def parse(wls):
    issue, traffic, lines = {}, {}, []
    for name in ('issue.json', 'traffic.json'):
        try:
            (issue if name == 'issue.json' else traffic).update(json.load(open(os.path.join(ROOT, 'profiles', name))))
        except (OSError, ValueError):
            pass
    for w in wls:
        try:
            meta = json.loads([l for l in open(os.path.join(OUT, 'cnt_%s.json' % w)) if l.startswith('{')][-1])
            rows = list(csv.reader(open(os.path.join(OUT, 'cnt_%s.csv' % w))))
        except (OSError, IndexError, ValueError):
            continue
        h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
        ci = {n: i for i, n in enumerate(rows[h])}
        per = {}
        for r in rows[h + 1:]:
            if len(r) <= ci['Metric Value']:
                continue
            k = r[ci['Kernel Name']]
            if 'at::' in k or 'elementwise' in k:        # torch's table-reset kernels are not part of the step
                continue
            d = per.setdefault(k, {'launches': set()})
            d['launches'].add(r[ci['ID']])
            m = r[ci['Metric Name']]
            d[m] = d.get(m, 0.0) + float(r[ci['Metric Value']].replace(',', ''))
        tot = {m: sum(d.get(m, 0.0) for d in per.values()) for m in METRICS.split(',')}
        units = meta['units']
        ipu = tot['smsp__inst_executed.sum'] / units
        dram = tot['dram__bytes_read.sum'] + tot['dram__bytes_write.sum']
        src = 'profiles/r2_counters.txt (ncu --metrics %s over one bench step)' % METRICS
        issue[w] = {'agents_per_gpu': meta['agents_per_gpu'], 'warp_instr_per_unit': ipu, 'units': units,
                    'kernel_ms_serialised': tot['gpu__time_duration.sum'] / 1e6, 'source': src}
        traffic[w] = {'agents_per_gpu': meta['agents_per_gpu'], 'bytes_per_launch': int(dram),
                      'note': 'dram__bytes_read.sum + dram__bytes_write.sum over all kernels of one bench step'}
        lines.append('== %s: %d agents/GPU, %.0f units per step; %.1f warp-instructions/unit, %.3f ms of kernels '
                     '(serialised, cold), DRAM %.1f MB read + %.1f MB written' %
                     (w, meta['agents_per_gpu'], units, ipu, tot['gpu__time_duration.sum'] / 1e6,
                      tot['dram__bytes_read.sum'] / 1e6, tot['dram__bytes_write.sum'] / 1e6))
        for k, d in sorted(per.items(), key=lambda kv: -kv[1].get('gpu__time_duration.sum', 0)):
            lines.append('   %-70s x%-3d %9.3f ms  %14.0f warp-instr  %8.1f MB DRAM' %
                         (k.replace('<unnamed>::', '')[:70], len(d['launches']), d.get('gpu__time_duration.sum', 0) / 1e6,
                          d.get('smsp__inst_executed.sum', 0),
                          (d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6))
    json.dump(issue, open(os.path.join(ROOT, 'profiles', 'issue.json'), 'w'), indent=1, sort_keys=True)
    json.dump(traffic, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1, sort_keys=True)
    open(os.path.join(ROOT, 'profiles', 'r2_counters.txt'), 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    mode = sys.argv[1] if len(sys.argv) > 1 else 'parse'
    wls = sys.argv[2:] or ALL
    (run if mode == 'run' else parse)(wls)
