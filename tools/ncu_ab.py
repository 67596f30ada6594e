#!/usr/bin/env python3
"""A/B table of two ncu SASS exports aggregated by device function: ncu_ab.py a.csv a.o b.csv b.o [kernel]"""
import subprocess, sys, os
HERE = os.path.dirname(os.path.abspath(__file__))
def table(csvf, obj, kern):
    out = subprocess.run([sys.executable, os.path.join(HERE, "ncu_by_function.py"), csvf, obj, kern], capture_output=True, text=True).stdout.splitlines()
    tot, ts = int(out[-1].split()[-1]), int(out[-1].split()[2])
    d = {}
    for l in out[1:-1]:
        name, rest = l[:30].strip(), l[30:].split()
        d[name[:22]] = (float(rest[0]) * ts / 100, float(rest[1]) * tot / 100, rest[2])
    return d, tot, ts
kern = sys.argv[5] if len(sys.argv) > 5 else "k_step"
a, ta, sa = table(sys.argv[1], sys.argv[2], kern)
b, tb, sb = table(sys.argv[3], sys.argv[4], kern)
print("total inst M: A %.0f B %.0f ; samples K: A %.0f B %.0f" % (ta / 1e6, tb / 1e6, sa / 1e3, sb / 1e3))
print("%-22s %9s %9s %7s | %9s %9s %7s" % ("function", "A instM", "B instM", "delta", "A smpK", "B smpK", "delta"))
for k in sorted(set(a) | set(b), key=lambda k: -(b.get(k, (0, 0))[0] - a.get(k, (0, 0))[0])):
    x, y = a.get(k, (0, 0, 0)), b.get(k, (0, 0, 0))
    if abs(y[1] - x[1]) > 2e6 or abs(y[0] - x[0]) > 1e3:
        print("%-22s %9.0f %9.0f %+7.0f | %9.1f %9.1f %+7.1f  lanes %s->%s" % (k, x[1] / 1e6, y[1] / 1e6, (y[1] - x[1]) / 1e6, x[0] / 1e3, y[0] / 1e3, (y[0] - x[0]) / 1e3, x[2], y[2]))
