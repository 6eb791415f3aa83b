import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(d['config']['workload'], 'steps/s %.2f e2e %.2f'%(d['value'],d['e2e']['value']), 'cg/step', d['config']['cg_iterations_per_step'], 'bicg', d['config']['bicgstab_iterations_last_step'])
    print('    ', {k:(v['ms'],v['gbs']) for k,v in d['kernels'].items()})
