"""Where do the membership questions that still reach the table come from?  (CPU simulation, measurement tooling.)

    python tools/bloom_sim.py [--scale=20] [--cap=256] [--samples=6000]
Samples node2vec contexts of a stationary walk on an R-MAT graph -- (t, v) a uniform stored edge, x a uniform neighbour of
v, x != t -- and classifies the question "x in adj(t)?" the way the walk kernel does: answered by the two triangle-Bloom
looks (the word of (t, v) asked about x, the word of (v, x) asked about t), or passed on because x is a common neighbour
(true member), because both words are saturated (both shorter rows exceed the cap: a hub triangle, the hub-pair filter's
case), or because of a Bloom false positive -- with one hash per element (shipped) and with two.
"""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from torch_random_walk_b200 import rmat  # noqa: E402


def arg(name, default):
    for a in sys.argv[1:]:
        if a.startswith(f"--{name}="):
            return a.split("=", 1)[1]
    return default


def main():
    scale, cap, n = int(arg("scale", "20")), int(arg("cap", "256")), int(arg("samples", "6000"))
    rp, ci = rmat.rmat_csr(scale, 16, device="cpu", seed=1)
    rp, ci = rp.numpy(), ci.numpy().astype(np.int64)
    rng = np.random.default_rng(0)
    m32 = np.uint64(0xFFFFFFFF)

    def h1(w):  # bloom_bit of member_table.cuh
        return int(((np.uint64(w) * np.uint64(0x9E3779B1)) & m32) >> np.uint64(27))

    def h2(w):
        return int(((np.uint64(w) * np.uint64(0x85EBCA6B)) & m32) >> np.uint64(27))

    def common(a, b):
        ra, rb = ci[rp[a]:rp[a + 1]], ci[rp[b]:rp[b + 1]]
        return None if min(ra.size, rb.size) > cap else np.intersect1d(ra, rb)

    es = rng.integers(0, ci.size, n)
    rows = np.searchsorted(rp, es, side="right") - 1
    need = member = saturated = fp1 = fp2 = 0
    for e, t in zip(es, rows):
        v = ci[e]
        x = ci[rp[v] + rng.integers(0, rp[v + 1] - rp[v])]
        if x == t:
            continue
        need += 1
        at = ci[rp[t]:rp[t + 1]]
        i = np.searchsorted(at, x)
        if i < at.size and at[i] == x:
            member += 1
            continue
        looks = [(common(t, v), x), (common(v, x), t)]
        if looks[0][0] is None and looks[1][0] is None:
            saturated += 1
            continue
        one = two = True
        for c, q in looks:
            if c is None:
                continue
            w1 = w2 = 0
            for y in c:
                w1 |= 1 << h1(y)
                w2 |= (1 << h1(y)) | (1 << h2(y))
            one = one and bool(w1 >> h1(q) & 1)
            two = two and bool((w2 >> h1(q) & 1) and (w2 >> h2(q) & 1))
        fp1 += one
        fp2 += two
    print(f"R-MAT scale {scale}, cap {cap}, {need} questions: true members {member / need:.3f}, both words saturated {saturated / need:.3f}, "
          f"Bloom false positives {fp1 / need:.3f} (one hash, shipped) / {fp2 / need:.3f} (two hashes)")


if __name__ == "__main__":
    main()
