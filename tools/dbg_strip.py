import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scribe_py_b200 as S
from oracle import ksg_oracle as o
dz, k = 3, 3
rng = np.random.default_rng(100 * dz + k); n = 700 + 37 * dz
z = rng.standard_normal((n, dz)); z[:, -1] = np.round(z[:, -1], 1)
x = 0.5 * z[:, :1] + rng.standard_normal((n, 1)); y = x + z[:, -1:] + 0.5 * rng.standard_normal((n, 1))
P = np.concatenate((x, y, z), axis=1)
eps, _ = o.knn_counts(P, [tuple(range(5))], k)
d = S.information_estimators.ksg_debug("cmi", x, y, z, k=k, path="strip")
bad = np.nonzero(d["eps"] != eps)[0]
print("mismatches", len(bad), "neg zeros", int((np.signbit(P[:, -1]) & (P[:, -1] == 0)).sum()))
# radix order: -0.0 before +0.0
key = P[:, -1].copy()
bits = key.view(np.int64)
order = np.lexsort((np.arange(n), np.where(bits < 0, ~bits, bits | (1 << 63)).astype(np.uint64)))
rank = np.empty(n, int); rank[order] = np.arange(n)
assert (np.diff(key[order]) >= 0).all()
for i in bad[:12]:
    dist = np.abs(P - P[i]).max(axis=1)
    js = np.argsort(dist, kind="stable")[:k + 1]
    miss = [j for j in js if dist[j] < d["eps"][i]]
    j = js[k]
    s0, sj = rank[i] // 64, rank[j] // 64
    # x-order position inside strips
    def xpos(a):
        st = order[(rank[a] // 64) * 64:(rank[a] // 64) * 64 + 64]
        return int(np.argsort(P[st, 0], kind="stable").tolist().index(list(st).index(a)))
    print("i", i, "strip", s0, "xpos", xpos(i), "| kth nbr j", j, "strip", sj, "xpos", xpos(j), "d", dist[j], "got", d["eps"][i], "zi", P[i, -1], "zj", P[j, -1], "dx", abs(P[i,0]-P[j,0]))
