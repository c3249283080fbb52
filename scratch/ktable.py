import json, sys
d=json.load(open(sys.argv[1]))
print('graph step ms', d['step_ms_graph'])
n = max(k['launches'] for k in d['kernels'])
tot=sum(k['ms']*k['launches'] for k in d['kernels'])/n
print('sum of kernel ms per step', tot)
for k in d['kernels']:
    w=k['work']
    rate = (w/(k['ms']*1e-3)/1e12 if k['kind']=='flops' else w/(k['ms']*1e-3)/1e9)
    print('%-24s %-44s %8.3f ms x%d share %5.1f%%  %8.1f %s' % (k['name'], k['args'][:10], k['ms'], k['launches']//n, 100*k['share'], rate, 'TF/s' if k['kind']=='flops' else 'GB/s'))
