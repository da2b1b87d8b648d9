"""Development helper (CPU only): numpy fp32 emulation of the recurrence-mode wave loops for one column of the
W-tier test scene — the reference's sequential accumulation, the rotation / three-term chain and the symmetric-pair
form, each against the exact (fp64) sum.   python tools/emulate_recurrence.py [waves] [grid_len]
Shows why the symmetric-pair form misses the parity bar: the reference itself is ~2e-4 sum(A) from the exact sum."""
import sys, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from common import Scene
from wave_tool_b200 import api, scenes
import common
f32 = np.float32
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 96
waves = scenes.seeded_waves(W, cfg=9)
hm = scenes.procedural_heightmap(64)
sc = Scene(waves, scenes.camera(), scenes.animation_params(5), L, hm)
st, fr = sc.frame()
bl = np.array(fr.bottomLeftGridPointInWorld[:], f32); br = np.array(fr.bottomRightGridPointInWorld[:], f32)
tl = np.array(fr.topLeftGridPointInWorld[:], f32); tr = np.array(fr.topRightGridPointInWorld[:], f32)
t = f32(fr.waveAnimationTimeInSeconds)
print("corners", bl, br, tl, tr, "t", t)
wv = sc.waves  # A, w, phi, Q, Dx, Dz
A, w, phi, Q, Dx, Dz = [wv[:, i].astype(f32) for i in range(6)]
cnt = max(W, 4)
Qi = (Q / (w * A * f32(cnt))).astype(f32)
du = f32(1.0) / f32(L - 1)
col = 58
u = f32(col) * du
def mix(a, b, x): return f32(f32(a * f32(f32(1) - x)) + f32(b * x))
m1 = np.array([mix(bl[0], br[0], u), mix(bl[2], br[2], u)], f32)
m2 = np.array([mix(tl[0], tr[0], u), mix(tl[2], tr[2], u)], f32)
e = ((m2 - m1) * du).astype(f32)
rows = np.arange(-1, 8)
def gridpos(r):
    v = f32(r) * du
    return np.array([mix(m1[0], m2[0], v), mix(m1[1], m2[1], v)], f32)
G = np.array([gridpos(r) for r in rows])
print("grid rows x", G[:, 0], "z", G[:, 1], "e", e)
# truth (fp64 on fp32 grid positions like the oracle does, but exact arithmetic)
def truth(r_idx):
    g = G[r_idx].astype(np.float64)
    k = w.astype(np.float64) * (Dx.astype(np.float64) * g[0] + Dz.astype(np.float64) * g[1]) + np.float64(f32(phi * t))
    return np.array([g[0] + np.sum(A * Qi * Dx * np.cos(k), dtype=np.float64), np.sum(A * np.sin(k), dtype=np.float64), g[1] + np.sum(A * Qi * Dz * np.cos(k), dtype=np.float64)])
def oracle32(r_idx):
    g = G[r_idx]
    d = (f32(Dx * g[0]) + f32(Dz * g[1])).astype(f32)
    k = (f32(w * d) + f32(phi * t)).astype(f32)
    c = np.cos(k.astype(np.float64)).astype(f32); s = np.sin(k.astype(np.float64)).astype(f32)
    return np.array([g[0] + np.sum(A * Qi * Dx * c, dtype=np.float64), np.sum(A * s, dtype=np.float64), g[1] + np.sum(A * Qi * Dz * c, dtype=np.float64)])
def mufu(x):   # sin/cos approx: angle quantised to fp32 turns, result fp32
    turns = (x.astype(f32) * f32(0.15915494309189535)).astype(f32).astype(np.float64)
    return np.sin(2 * np.pi * turns).astype(f32), np.cos(2 * np.pi * turns).astype(f32)
def seedphase(g):
    wDx = (w * Dx).astype(f32); wDz = (w * Dz).astype(f32)
    ph = np.remainder(np.float64(1) * f32(phi * t).astype(np.float64) + np.pi, 2 * np.pi) - np.pi
    k = (wDx.astype(np.float64) * g[0] + (wDz.astype(np.float64) * g[1] + ph.astype(f32))).astype(f32)  # 2 fma: approx
    n = np.rint(k.astype(np.float64) / (2 * np.pi))
    r = (k.astype(np.float64) - n * 2 * np.pi).astype(f32)
    return r, wDx, wDz
def prod():
    r, wDx, wDz = seedphase(G[0])
    s, c = mufu(r)
    dk = (wDx.astype(np.float64) * e[0] + f32(wDz * e[1]).astype(np.float64)).astype(f32)
    sd, cd = mufu(dk)
    out = []
    for j in range(9):
        if j > 0:
            c2 = (c * cd - s * sd).astype(f32); s = (s * cd + c * sd).astype(f32); c = c2
        out.append(np.array([G[j][0] + np.sum(A * Qi * Dx * c, dtype=np.float64), np.sum(A * s, dtype=np.float64), G[j][1] + np.sum(A * Qi * Dz * c, dtype=np.float64)]))
    return out, dk
def sym():
    H = 4
    r, wDx, wDz = seedphase(G[H])
    s, c = mufu(r)
    dk = (wDx.astype(np.float64) * e[0] + f32(wDz * e[1]).astype(np.float64)).astype(f32)
    S = [None] * (H + 1); C = [None] * (H + 1)
    S[1], C[1] = mufu(dk)
    for j in range(2, H + 1):
        if j % 2 == 0:
            t2 = (S[j // 2] + S[j // 2]).astype(f32)
            C[j] = (f32(1) - t2 * S[j // 2]).astype(f32); S[j] = (t2 * C[j // 2]).astype(f32)
        else:
            C[j] = (C[j - 1] * C[1] - S[j - 1] * S[1]).astype(f32); S[j] = (S[j - 1] * C[1] + C[j - 1] * S[1]).astype(f32)
    out = [None] * 9
    cX = (A * Qi * Dx).astype(f32); cZ = (A * Qi * Dz).astype(f32)
    out[H] = np.array([G[H][0] + np.sum(cX * c, dtype=np.float64), np.sum(A * s, dtype=np.float64), G[H][1] + np.sum(cZ * c, dtype=np.float64)])
    for j in range(1, H + 1):
        ax = np.sum((cX * c).astype(f32) * C[j], dtype=np.float64); bx = np.sum((cX * s).astype(f32) * S[j], dtype=np.float64)
        az = np.sum((cZ * c).astype(f32) * C[j], dtype=np.float64); bz = np.sum((cZ * s).astype(f32) * S[j], dtype=np.float64)
        ay = np.sum((A * s).astype(f32) * C[j], dtype=np.float64); by = np.sum((A * c).astype(f32) * S[j], dtype=np.float64)
        out[H + j] = np.array([G[H + j][0] + ax - bx, ay + by, G[H + j][1] + az - bz])
        out[H - j] = np.array([G[H - j][0] + ax + bx, ay - by, G[H - j][1] + az + bz])
    return out, dk
P, dk = prod(); Sy, _ = sym()
print("dk range", np.abs(dk).min(), np.abs(dk).max())
tot = float(A.sum())
for j in range(9):
    T = truth(j); O = oracle32(j)
    print(rows[j], "oracle-truth %.2e  prod-oracle %.2e  sym-oracle %.2e   prod-truth %.2e sym-truth %.2e" % (
        np.abs(O - T).max() / tot, np.abs(P[j] - O).max() / tot, np.abs(Sy[j] - O).max() / tot, np.abs(P[j] - T).max() / tot, np.abs(Sy[j] - T).max() / tot))
print("---- sequential fp32 accumulation like the shader")
def oracle_seq(r_idx):
    g = G[r_idx]
    d = (f32(Dx * g[0]) + f32(Dz * g[1])).astype(f32)
    k = (f32(w * d) + f32(phi * t)).astype(f32)
    c = np.cos(k.astype(np.float64)).astype(f32); s = np.sin(k.astype(np.float64)).astype(f32)
    x, y, z = g[0], f32(0), g[1]
    tx = (f32(A * Qi) * Dx).astype(f32)
    for i in range(W):
        x = f32(x + f32(f32(f32(Qi[i] * A[i]) * Dx[i]) * c[i])); y = f32(y + f32(A[i] * s[i])); z = f32(z + f32(f32(f32(Qi[i] * A[i]) * Dz[i]) * c[i]))
    return np.array([x, y, z], np.float64)
for j in range(9):
    T = truth(j); O = oracle_seq(j)
    print(rows[j], "oracleseq-truth %.2e   sym-oracleseq %.2e  prod(f64 sums)-oracleseq %.2e" % (np.abs(O - T).max() / tot, np.abs(Sy[j] - O).max() / tot, np.abs(P[j] - O).max() / tot), (O - T) / tot)
