"""Second, independent restatement of the reference math in plain Python/float64 for SMALL images only
(pure loops).  Written from SURVEY.md Appendix A, not from oracle/svgf_oracle.cpp, so that a transcription
slip in either shows up as a disagreement.  Storage is modelled as float32 ("F32 mode"): values are rounded
to float32 where the reference stores them, everything else is float64 — agreement with the C++ oracle is
therefore expected to ~1e-6, not bitwise."""
import math

import numpy as np

KW = [1.0, 2.0 / 3.0, 1.0 / 6.0]


def f32(x):
    return float(np.float32(x))


def lum(c):
    return 0.2126 * c[0] + 0.7152 * c[1] + 0.0722 * c[2]


def depth(motion, x, y):
    z, dz = float(motion[y, x, 2]), float(motion[y, x, 3])
    return (1e30, 0.0) if z == 0.0 else (z, dz)


def nrm(normal_f, x, y):
    return [float(v) for v in normal_f[y, x, :3]]


def clamp01(v):
    return min(max(v, 0.0), 1.0)


def weight(zc, zq, phiZ, nc, nq, phiN, lc, lq, phiL):
    d = nc[0] * nq[0] + nc[1] * nq[1] + nc[2] * nq[2]
    d = 0.0 if d != d else min(max(d, 0.0), 1.0)
    wN = d ** phiN if not (d == 0.0 and phiN == 0.0) else 1.0
    wZ = 0.0 if phiZ == 0.0 else abs(zc - zq) / phiZ
    wL = abs(lc - lq) / phiL
    return math.exp(-max(wL, 0.0) - max(wZ, 0.0)) * wN


def temporal(p, cur, prev, prev_col, cur_col, hist, prev_mom):
    """cur/prev: dicts with 'normal' (float decoded [H,W,4]), 'inst' (int [H,W]), 'motion'.  Returns (col, hist, mom)."""
    H, W = hist.shape
    out_c = np.zeros((H, W, 4)); out_h = np.zeros((H, W), np.uint8); out_m = np.zeros((H, W, 2))
    for y in range(H):
        for x in range(W):
            c = [clamp01(float(v)) for v in cur_col[y, x, :3]]
            mv = cur["motion"][y, x]

            def consistent(qx, qy):
                if not (0 <= qx < W and 0 <= qy < H):
                    return False
                dzv = abs(depth(prev["motion"], qx, qy)[0] - depth(cur["motion"], x, y)[0])
                if getattr(p, "depth_test_mode", 0) == 1:   # SVGF_DEPTH_TEST_RELATIVE, src/Filter.cuh:241
                    if f32(f32(dzv) / f32(f32(depth(cur["motion"], x, y)[1]) + f32(1e-2))) > p.depth_threshold:
                        return False
                elif dzv > p.depth_threshold:
                    return False
                if p.mesh_id_mode == 0 and int(cur["inst"][y, x]) != int(prev["inst"][qy, qx]):
                    return False
                a, b = nrm(cur["normal"], x, y), nrm(prev["normal"], qx, qy)
                return not ((a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) < p.normal_threshold)

            if getattr(p, "reproj_mode", 0) == 1:      # SVGF_REPROJ_BILINEAR (include/svgf.h), float64 here
                fx, fy = f32(f32(x) + f32(mv[0])), f32(f32(y) + f32(mv[1]))
                x0, y0 = math.floor(fx), math.floor(fy)
                tx, ty = fx - x0, fy - y0
                ws, hs, cs, ms = 0.0, 0.0, [0.0] * 3, [0.0] * 2
                for j, wy in ((0, 1 - ty), (1, ty)):
                    for k, wx in ((0, 1 - tx), (1, tx)):
                        qx, qy = x0 + k, y0 + j
                        if consistent(qx, qy):
                            w = wx * wy
                            ws += w
                            hs += w * int(hist[qy, qx])
                            cs = [cs[n] + w * clamp01(float(prev_col[qy, qx, n])) for n in range(3)]
                            ms = [ms[n] + w * float(prev_mom[qy, qx, n]) for n in range(2)]
                ok = ws >= 0.01
                if ok:
                    pc, pm = [v / ws for v in cs], [v / ws for v in ms]
                    h = min(p.history_cap, int(hs / ws + 0.5) + 1)
                    alpha = f32(1.0 / h)
            else:
                qx, qy = x + int(mv[0]), y + int(mv[1])  # int() truncates toward zero
                ok = consistent(qx, qy)
                if ok:
                    pc = [clamp01(float(v)) for v in prev_col[qy, qx, :3]]
                    h = min(p.history_cap, int(hist[qy, qx]) + 1)
                    pm = [float(v) for v in prev_mom[qy, qx]]
                    alpha = f32(1.0 / h)
            if not ok:
                pc, pm, h, alpha = [0, 0, 0], [0, 0], 1, 1.0
            L = lum(c)
            m = [pm[0] * (1 - alpha) + L * alpha, pm[1] * (1 - alpha) + L * L * alpha]
            var = max(0.0, m[1] - m[0] * m[0])
            col = [pc[k] * (1 - alpha) + c[k] * alpha for k in range(3)]
            out_h[y, x] = h
            out_c[y, x] = [clamp01(v) for v in col] + [clamp01(var)]
            out_m[y, x] = m
    return out_c, out_h, out_m


def variance(p, g, col, mom, hist):
    H, W = hist.shape
    out = np.zeros((H, W, 4))
    for y in range(H):
        for x in range(W):
            h = float(hist[y, x])
            if h >= 4:
                out[y, x] = col[y, x]
                continue
            cc = [float(v) for v in col[y, x]]
            lc = lum(cc)
            zc, dzc = depth(g["motion"], x, y)
            nc = nrm(g["normal"], x, y)
            phiZ0 = f32(max(dzc, 1e-8) * 3.0) * p.phi_depth
            S, C, M = 0.0, [0.0] * 3, [0.0] * 2
            for yy in range(-3, 4):
                for xx in range(-3, 4):
                    px, py = x + xx, y + yy
                    if not (0 <= px < W and 0 <= py < H):
                        continue
                    cq = [float(v) for v in col[py, px]]
                    mq = [float(v) for v in mom[py, px]]
                    w = weight(zc, depth(g["motion"], px, py)[0], phiZ0 * math.sqrt(xx * xx + yy * yy), nc,
                               nrm(g["normal"], px, py), p.phi_normal, lc, lum(cq), p.phi_colour)
                    S += w
                    C = [C[k] + cq[k] * w for k in range(3)]
                    M = [M[k] + mq[k] * w for k in range(2)]
            S = max(S, 1e-6)
            C = [v / S for v in C]
            M = [v / S for v in M]
            var = (M[1] - M[0] * M[0]) * (4.0 / h) if h > 0 else float("inf")
            out[y, x] = C + [var]
    return out


def atrous(p, g, inp, level, hist_colour=None):
    H, W = inp.shape[:2]
    step = 1 << level
    out = np.zeros((H, W, 4))
    hc = None if hist_colour is None else hist_colour.copy()
    for y in range(H):
        for x in range(W):
            c = [clamp01(float(v)) for v in inp[y, x]]
            lc = lum(c)
            var = c[3]
            if getattr(p, "variance_prefilter", 0) == 1:     # SVGF_VARIANCE_PREFILTER_GAUSS3: (1 2 1; 2 4 2; 1 2 1) / 16, clamped coordinates
                var = 0.0
                for dy, wy in ((-1, 0.25), (0, 0.5), (1, 0.25)):
                    for dx, wx in ((-1, 0.25), (0, 0.5), (1, 0.25)):
                        var += wy * wx * clamp01(float(inp[min(max(y + dy, 0), H - 1), min(max(x + dx, 0), W - 1), 3]))
            zc, dzc = depth(g["motion"], x, y)
            if zc == 1e30:
                out[y, x] = c
                continue
            nc = nrm(g["normal"], x, y)
            phiL = p.phi_colour * math.sqrt(max(0.0, f32(f32(1e-10) + f32(var))))
            phiZ = f32(max(dzc, 1e-6)) * step * p.phi_depth
            S, A = 1.0, list(c)
            for yy in range(-2, 3):
                for xx in range(-2, 3):
                    px, py = x + xx * step, y + yy * step
                    if not (0 <= px < W and 0 <= py < H) or (xx == 0 and yy == 0):
                        continue
                    k = f32(KW[abs(xx)]) * f32(KW[abs(yy)])
                    cq = [clamp01(float(v)) for v in inp[py, px]]
                    w = weight(zc, depth(g["motion"], px, py)[0], phiZ * math.sqrt(xx * xx + yy * yy), nc,
                               nrm(g["normal"], px, py), p.phi_normal, lc, lum(cq), phiL) * k
                    S += w
                    A = [A[0] + w * cq[0], A[1] + w * cq[1], A[2] + w * cq[2], A[3] + w * w * cq[3]]
            o = [A[0] / S, A[1] / S, A[2] / S, A[3] / (S * S)]
            out[y, x] = o
            if level == 0 and hc is not None:
                hc[y, x] = o
    return out, hc
