import json, sys
for f in sys.argv[1:]:
    d=json.load(open(f))
    rows=d['kernels']
    n=max(r['launches'] for r in rows)
    s=sum(r['ms']*r['launches'] for r in rows)/n
    print(f, 'graph step ms', d['step_ms_graph'], ' sum of entries per step', s)
    for r in sorted(rows, key=lambda r:-r['ms']*r['launches'])[:60]:
        print('   %-32s %-44s %7.1f us x%.1f' % (r['name'], r['args'][:8], r['ms']*1e3, r['launches']/n))
