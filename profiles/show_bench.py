import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, 'ERR', e); continue
    r=d['roofline']
    print(f, 'value %.1f'%d['value'], 'frac %.3f'%(r['frac'] or 0), 'avg_launch_ms', r.get('avg_launch_ms'), d['solver']['path'], d['solver']['iterations_mean'])
    for k,v in r['other_kernels'].items():
        print('    ', k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
