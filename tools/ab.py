#!/usr/bin/env python
"""A/B of kernel tuning builds: python tools/ab.py [lib.so ...]  (default: the in-tree library, then every tunelibs/*.so).
Each library runs `tools/tune.py ab` in its own process (SCRIBE_B200_LIB); estimator-kernel times and checksums side by side."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:] or [os.path.join(ROOT, "scribe_py_b200", "lib", "libscribe_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "tunelibs", "*.so")))
rows = {}
for lib in libs:
    env = dict(os.environ, SCRIBE_B200_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tune.py"), "ab"], env=env, capture_output=True, text=True)
    name = os.path.basename(lib).replace("libscribe_b200", "").replace(".so", "").strip("_") or "default"
    if out.returncode != 0:
        print("LIB %s FAILED:\n%s" % (name, out.stderr[-2000:]), flush=True)
        continue
    for ln in out.stdout.splitlines():
        if ln.startswith("AB "):
            f = ln.split()
            rows.setdefault(f[1], []).append((name, f[3], f[4], f[5], f[6]))
for wl, r in rows.items():
    print("== %s" % wl)
    for name, est, tot, vis, chk in r:
        print("   %-12s %-18s %-20s %-16s %s" % (name, est, tot, vis, chk))
