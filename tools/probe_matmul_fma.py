"""Which arithmetic does numpy matmul (OpenBLAS) use for the small products of the build path?
Compares A @ X with fused / unfused, left-to-right / right-to-left restatements (exact rational FMA).
Outcome in the build container: fma_ltr reproduces numpy on every element (oracle/build_oracle.c: dot3)."""
import numpy as np
from fractions import Fraction as F
rng = np.random.default_rng(0)
def fma(a,b,c): return float(F(a)*F(b)+F(c))
def variants(row, x):
    n = len(row)
    out = {}
    # non-fma left to right
    acc = row[0]*x[0]
    for k in range(1,n): acc = acc + row[k]*x[k]
    out['nofma_ltr'] = acc
    acc = row[0]*x[0]
    for k in range(1,n): acc = fma(row[k],x[k],acc)
    out['fma_ltr'] = acc
    acc = row[n-1]*x[n-1]
    for k in range(n-2,-1,-1): acc = fma(row[k],x[k],acc)
    out['fma_rtl'] = acc
    acc = row[n-1]*x[n-1]
    for k in range(n-2,-1,-1): acc = acc + row[k]*x[k]
    out['nofma_rtl'] = acc
    if n==4:
        out['pair_nofma'] = (row[0]*x[0]+row[1]*x[1])+(row[2]*x[2]+row[3]*x[3])
        out['pair_fma'] = fma(row[1],x[1],row[0]*x[0]) + fma(row[3],x[3],row[2]*x[2])
    return out
def probe(name, A, X):
    Y = A @ X
    counts = {}
    tot = 0
    for j in range(X.shape[1]):
        for i in range(A.shape[0]):
            v = variants([float(a) for a in A[i]], [float(x) for x in X[:,j]])
            tot += 1
            for k,val in v.items():
                counts[k] = counts.get(k,0) + (val == Y[i,j])
    print(name, tot, counts)
K = np.array([[54.,0,54],[0,54,36],[0,0,1]]); Kinv = np.linalg.inv(K)
Kr = rng.standard_normal((3,3))
for N in (1, 2, 7, 64, 500, 3072):
    X = rng.standard_normal((3,N))
    probe(f"3x3 rand @ 3x{N}", Kr, X)
T = rng.standard_normal((4,4)); T[3]=[0,0,0,1]
for N in (1, 7, 64, 500, 3072):
    X = np.vstack([rng.standard_normal((3,N)), np.ones((1,N))])
    probe(f"4x4 @ 4x{N}", T, X)
