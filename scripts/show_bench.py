import json, sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, 'ERR', e); 
        try: print(open(f.replace('.json','.err')).read()[-1500:])
        except Exception: pass
        continue
    print('==',f, 'value=%.1f'%d['value'], 'ms/step=%.3f'%d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'), 'launches', d.get('gpu_launches'), d.get('config',{}).get('cuda_graph'))
    if 'roofline' in d: print('  roofline', {k:d['roofline'][k] for k in ('kernel','achieved','peak','frac','us_per_launch')})
    if 'kernel_shares' in d:
        tot=0
        for k,v in d['kernel_shares'].items():
            print('   %-28s calls/step %.1f us/call %8.1f share %s'%(k,v['calls_per_step'],v['us_per_call'],v['share'])); tot+=v['calls_per_step']*v['us_per_call']
        print('   sum of kernel time per step: %.1f us'%tot)
    if 'clocks' in d: print('  clocks', d['clocks'])
    if 'homography_adaptation' in d:
        a=d['homography_adaptation']; print('  adapt value %.0f img/s  ms/step %.3f  e2e %s  cpu %s' % (a['value'], a.get('ms_per_step',0), a.get('e2e',{}).get('value'), a.get('cpu_baseline',{}).get('value')))
        if 'roofline' in a: print('   adapt roofline', {k:a['roofline'][k] for k in ('kernel','achieved','frac','us_per_launch')}, a['roofline']['note'][-40:])
        for k,v in a.get('kernel_shares',{}).items(): print('     %-28s calls/step %.1f us/call %8.1f share %.3f'%(k,v['calls_per_step'],v['us_per_call'],v['share']))
    if 'cpu_baseline' in d: print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
    if 'with_semantic_head' in d: print('  semantic', d['with_semantic_head'])
    if 'variants' in d: print('  variants', {k: (round(v['ms_per_step'], 4) if isinstance(v, dict) and 'ms_per_step' in v else v) for k, v in d['variants'].items()} if isinstance(d['variants'], dict) else d['variants'])
