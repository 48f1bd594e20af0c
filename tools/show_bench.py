"""Pretty-print one bench.py JSON line (headline + workloads)."""
import json
import sys


def show(r, name):
    rf = r['roofline']
    print("%-18s value %9.1f e2e %9.1f ms/step %8.3f e2e_ms %8.3f frac %.3f share %s h2d %s GB/s launches %s" % (
        name, r['value'], r['e2e']['value'], r['ms_per_step'], r['e2e'].get('ms_per_step', 0), rf['frac'],
        rf.get('kernel_share_of_step'), r['e2e'].get('h2d_copy_gbs'), r.get('gpu_launches')))
    if '-v' in sys.argv:
        for k in rf.get('hbm_kernels') or []:
            print("      hbm", k)
        for k in rf.get('top_ops') or []:
            print("      top", k)


for path in [a for a in sys.argv[1:] if not a.startswith('-')]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    show(d, d['config']['workload'])
    for w in d.get('workloads', []):
        show(w, w['workload'])
