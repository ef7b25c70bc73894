#!/usr/bin/env python
"""Generates nonlocal_image_edit_b200/csrc/lab_tables.inc: the integer look-up tables of OpenCV's 8-bit
BGR<->Lab conversion (the colour conversion NLEFilter::enhance / trainForEnhancement perform around the hot
path: cv::cvtColor(..., COLOR_BGR2Lab / COLOR_Lab2BGR) on CV_8UC3, reference filter.cpp:422-426,438-440,463-466).

OpenCV is an un-vendored dependency of the reference (CMakeLists.txt:34, unpinned); its 8-bit Lab path is fixed-point
arithmetic over small tables (imgproc/src/color_lab.cpp: RGB2Lab_b, Lab2RGBinteger).  This script restates those
tables from their published formulas, then VERIFIES the complete integer pipeline against cv2.cvtColor of the installed
opencv-python on all 2^24 BGR triples and all 2^24 Lab triples before writing anything.  One entry of the cube-root
table (index 324, an exact .5 tie in binary32) is calibrated to what cv2 produces; everything else follows the formulas.
Run here (needs cv2); the .inc it writes is committed and travels to the GPU box."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "nonlocal_image_edit_b200", "csrc", "lab_tables.inc")

GAMMA_SHIFT, LAB_SHIFT, LAB_SHIFT2 = 3, 12, 15
BASE, INV_TAB = 1 << 14, 4096
MIN_AB = -8145
D65 = [0.950456, 1.0, 1.088754]
RGB2XYZ = [0.412453, 0.357580, 0.180423, 0.212671, 0.715160, 0.072169, 0.019334, 0.119193, 0.950227]
XYZ2RGB = [3.240479, -1.53715, -0.498535, -0.969256, 1.875991, 0.041556, 0.055648, -0.204043, 1.057311]


def tables():
    f32 = np.float32
    i = np.arange(256)
    x = (i * f32(1 / 255.0)).astype(f32)
    g = np.where(x <= f32(0.04045), x * f32(1 / 12.92),
                 np.power((x.astype(np.float64) + 0.055) * (1 / 1.055), 2.4).astype(f32)).astype(f32)
    gamma = np.clip(np.rint(f32(255.0 * 8) * g), 0, 65535).astype(np.int64)
    n = 256 * 3 // 2 * 8
    xi = (np.arange(n) * f32(1.0 / (255.0 * 8))).astype(f32)
    cb = np.where(xi < f32(0.008856), xi * f32(7.787) + f32(0.13793103448275862), np.cbrt(xi.astype(np.float64)).astype(f32))
    cbrt = np.clip(np.rint(f32(1 << LAB_SHIFT2) * cb), 0, 65535).astype(np.int64)
    cbrt[324] -= 1            # 17745.5 in binary32: cv2's cvCbrt lands just below the tie (calibrated, verified below)
    ytab = np.zeros(256, np.int64)
    fytab = np.zeros(256, np.int64)
    for L in range(256):
        if L <= 20:
            y = int(np.rint(f32(L * BASE * 20 * 9) / f32(17 * 29 * 29 * 29)))
            ify = int(np.rint(f32(BASE) * (f32(16) / f32(116) + f32(L * 5) / f32(3 * 17 * 29))))
        else:
            fy = f32(f32(L * 100 * BASE) / f32(255 * 116) + f32(16 * BASE) / f32(116))
            ify = int(np.rint(fy))
            y = int(np.rint(f32(f32(f32(fy * fy) * fy) / f32(BASE * BASE))))
        ytab[L], fytab[L] = y, ify
    xg = (np.arange(INV_TAB, dtype=f32) * f32(1.0 / INV_TAB)).astype(f32)
    ig = np.where(xg <= f32(0.0031308), xg * f32(12.92),
                  (f32(1.055) * np.power(xg.astype(np.float64), 1 / 2.4).astype(f32) - f32(0.055)).astype(f32))
    invgamma = np.rint(f32(255) * ig.astype(f32)).astype(np.int64)
    cf = [int(np.rint(RGB2XYZ[r * 3 + c] * (1 << LAB_SHIFT) / D65[r])) for r in range(3) for c in range(3)]
    ci = [int(np.rint((1 << LAB_SHIFT) * XYZ2RGB[r * 3 + c] * D65[c])) for r in range(3) for c in range(3)]
    return gamma, cbrt, ytab, fytab, invgamma, cf, ci


def descale(x, n):
    return (x + (1 << (n - 1))) >> n


def ab_to_xz(i):
    """abToXZ_b of color_lab.cpp, pure integer (C truncating division)."""
    lo = np.sign(i * 108) * (np.abs(i * 108) // 841) - (BASE * 16 // 116 * 108 // 841)
    hi = (i * i // BASE) * i // BASE
    return np.where(i <= 3390, lo, hi)


def bgr2lab(bgr, T):
    gamma, cbrt, _, _, _, C, _ = T
    B, G, R = gamma[bgr[..., 0]], gamma[bgr[..., 1]], gamma[bgr[..., 2]]
    fX = cbrt[descale(R * C[0] + G * C[1] + B * C[2], LAB_SHIFT)]
    fY = cbrt[descale(R * C[3] + G * C[4] + B * C[5], LAB_SHIFT)]
    fZ = cbrt[descale(R * C[6] + G * C[7] + B * C[8], LAB_SHIFT)]
    Lscale = (116 * 255 + 50) // 100
    Lshift = -((16 * 255 * (1 << LAB_SHIFT2) + 50) // 100)
    L = descale(Lscale * fY + Lshift, LAB_SHIFT2)
    a = descale(500 * (fX - fY) + 128 * (1 << LAB_SHIFT2), LAB_SHIFT2)
    b = descale(200 * (fY - fZ) + 128 * (1 << LAB_SHIFT2), LAB_SHIFT2)
    return np.stack([np.clip(L, 0, 255), np.clip(a, 0, 255), np.clip(b, 0, 255)], -1).astype(np.uint8)


def lab2bgr(lab, T):
    _, _, ytab, fytab, invgamma, _, C = T
    L, a, b = (lab[..., k].astype(np.int64) for k in range(3))
    y, ify = ytab[L], fytab[L]
    adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * BASE // 500
    bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * BASE // 200 + 1
    x, z = ab_to_xz(ify + adiv), ab_to_xz(ify - bdiv)
    out = []
    for r in range(3):
        v = descale(C[r * 3] * x + C[r * 3 + 1] * y + C[r * 3 + 2] * z, 14)
        out.append(invgamma[np.clip(v, 0, INV_TAB - 1)])
    return np.stack([out[2], out[1], out[0]], -1).astype(np.uint8)


def cube(first):
    for v0 in range(0, 256, 8):
        a, b, c = np.meshgrid(np.arange(v0, v0 + 8), np.arange(256), np.arange(256), indexing="ij")
        order = [a, b, c] if first else [c, b, a]
        yield np.stack(order, -1).astype(np.uint8).reshape(-1, 1, 3)


def verify(T):
    bad_f = bad_i = 0
    for blk in cube(False):
        bad_f += int((cv2.cvtColor(blk, cv2.COLOR_BGR2Lab) != bgr2lab(blk, T)).sum())
    for blk in cube(True):
        bad_i += int((cv2.cvtColor(blk, cv2.COLOR_Lab2BGR) != lab2bgr(blk, T)).sum())
    return bad_f, bad_i


def carr(name, ctype, v, per=16):
    rows = [", ".join(str(int(x)) for x in v[i:i + per]) for i in range(0, len(v), per)]
    return f"NLE_LAB_TAB {ctype} {name}[{len(v)}] = {{\n    " + ",\n    ".join(rows) + "\n};\n"


if __name__ == "__main__":
    T = tables()
    bf, bi = verify(T)
    print(f"cv2 {cv2.__version__}: BGR2Lab mismatching bytes {bf} / {3 << 24}, Lab2BGR mismatching bytes {bi} / {3 << 24}")
    if bf or bi:
        sys.exit("tables do not reproduce cv2 -- not writing")
    gamma, cbrt, ytab, fytab, invgamma, cf, ci = T
    with open(OUT, "w") as f:
        f.write("// Generated by scripts/make_lab_tables.py -- do not edit.  Integer tables of OpenCV's 8-bit BGR<->Lab\n"
                f"// (imgproc color_lab.cpp: RGB2Lab_b, Lab2RGBinteger); verified against cv2 {cv2.__version__} on all 2^24 BGR and all\n"
                "// 2^24 Lab triples (0 mismatching bytes in either direction).  The includer defines NLE_LAB_TAB (storage qualifiers).\n")
        f.write(carr("kLabGammaTab", "unsigned short", gamma))
        f.write(carr("kLabCbrtTab", "unsigned short", cbrt))
        f.write(carr("kLabYTab", "unsigned short", ytab))
        f.write(carr("kLabFyTab", "unsigned short", fytab))
        f.write(carr("kLabInvGammaTab", "unsigned char", invgamma))
        f.write(carr("kLabFwdCoef", "int", cf))
        f.write(carr("kLabInvCoef", "int", ci))
    print("wrote", OUT)
