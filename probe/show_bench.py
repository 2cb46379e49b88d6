"""Print the figures of a bench.py JSON line (last '{' line of the given file)."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('headline', d['config']['workload'][:3], round(d['value'], 3), 'TF', round(d['ms_per_step'], 3), 'ms', 'plan', d.get('details', {}).get('plan'),
      'frac', round(d['roofline']['frac'], 4), 'e2e', d.get('e2e', {}).get('value'), d.get('kernel'))
for k, v in d.get('per_config', {}).items():
    vr = v.get('vs_reference') or {}
    print(f"{k:20s} {v['value']:9.2f} {v.get('ms_per_step', v.get('ms')):9.4f} ms  frac {v['roofline']['frac']:.3f}  plan {v.get('plan')}  ev {v.get('event_time', {}).get('ms')}"
          f"  ref {vr.get('reference_ms', vr.get('reference_gbs'))}  ratio {vr.get('ratio')}  step {vr.get('ratio_step')}  {v.get('kernel')}")
    if 'best_ms_per_plan' in v:
        print('    ', v['best_ms_per_plan'])
