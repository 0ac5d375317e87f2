import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
        print(f, 'value %.0f ms/step %.4f e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']), 'phases', {k:round(v,4) for k,v in d['phases_ms'].items() if len(k)==1}, 'roof', round(d['step_roofline']['step']['frac'],3),
              'eval', round(d.get('eval',{}).get('users_per_sec',0)), d.get('eval',{}).get('error'))
    except Exception as e: print(f, 'ERR', e)
