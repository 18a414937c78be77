#!/bin/bash
# quick GPU check: bench of the configs given as $1 (comma list) with extra args $2.., one line per config
only=$1; shift
python bench.py --only $only --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        for c in d['all_configs']:
            r=c.get('roofline') or {}
            print(c['workload'], 'value %.4g' % c['value'], 'ms/step %.4f' % c['ms_per_step'], 'kernel_ms %.4f' % c['kernel_ms'], 'frac %.3f' % (r.get('frac') or 0), 'e2e %.4g' % ((c.get('e2e') or {}).get('value') or 0), 'lat_us', c.get('latency_us_per_call'))
    elif 'rror' in l: print(l.rstrip())
"
