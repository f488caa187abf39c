import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print('%s: value %.3e e2e %.3e k_step %.1f us frac %.3f gmax %.1f us whole %.3f' % (
            f, d['value'], d['e2e']['value'], r.get('kernel_us', 0), r['frac'], r.get('gmax_kernel_us', 0), r.get('whole_step_frac', 0)))
    except Exception as e:
        print(f, 'FAILED', e)
