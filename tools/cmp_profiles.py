#!/usr/bin/env python
"""Compare two per-op CUDA-event profiles written by bench.py --profile-out (conv launches only)."""
import json, sys
a = json.load(open(sys.argv[1])); b = json.load(open(sys.argv[2]))
ta = tb = 0
for x, y in zip(a, b):
    if x['kind'] == 'conv_igemm' and x['desc']:
        ta += x['ms']; tb += y['ms']
        print(f"{x['ms']*1e3:7.1f} {y['ms']*1e3:7.1f} us  {x['flops']/x['ms']/1e9:6.0f} {y['flops']/y['ms']/1e9:6.0f} TF/s  {x['desc']}  ->  {y['desc'].split('tiles')[1]}")
print(ta, tb)
for k in ('groupnorm', 'other', 'attention'):
    print(k, sum(x['ms'] for x in a if x['kind'] == k), sum(x['ms'] for x in b if x['kind'] == k))
