"""CPU: the per-pixel arithmetic of csrc/rays.cu (struct RayCam .. ray_for_pixel), compiled for the HOST with g++ -- the
round-to-nearest intrinsics mapped to un-contracted IEEE operations and fma() -- must reproduce the reference's fixtures
(tests/golden/rays_*.npz, written by camera_util.py itself): ray_mask bit-exact and o, d, near, far bitwise identical.
This checks the kernel's rounding sequence without a GPU; it is a test harness around the device function's text, not
a CPU path of the product (nothing in occnerf_b200/ can reach it).
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "oracle", "_build")

PRELUDE = r"""
#include <math.h>
#include <stdint.h>
#define __device__
#define __forceinline__ inline
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
"""

DRIVER = r"""
extern "C" void emulate(const double *kinv, int k_f32, const double *R, const double *T, const double *lo, const double *hi,
                        int H, int W, float *rays, uint8_t *mask) {
    RayCam c;
    for (int a = 0; a < 9; ++a) { c.kinv[a] = kinv[a]; c.R[a] = R[a]; }
    for (int a = 0; a < 3; ++a) {
        c.T[a] = T[a]; c.lo[a] = lo[a] + -0.01; c.hi[a] = hi[a] + 0.01;
        if (k_f32 == 2) {
            float acc = (float)R[a] * (float)T[0]; acc = fmaf((float)R[3 + a], (float)T[1], acc); acc = fmaf((float)R[6 + a], (float)T[2], acc);
            c.o[a] = (double)-acc;
        } else {
            double acc = R[a] * T[0]; acc = fma(R[3 + a], T[1], acc); acc = fma(R[6 + a], T[2], acc);
            c.o[a] = -acc;
        }
    }
    c.H = H; c.W = W; c.k_f32 = k_f32 ? 1 : 0; c.all_f32 = k_f32 == 2;
    for (int p = 0; p < H * W; ++p) {
        RayOut r = ray_for_pixel(c, p);
        mask[p] = r.hit;
        float *q = rays + 8 * (long)p;
        for (int a = 0; a < 3; ++a) { q[a] = (float)c.o[a]; q[3 + a] = (float)r.d[a]; }
        q[6] = r.near; q[7] = r.far;
    }
}
"""


def _build():
    src = open(os.path.join(ROOT, "occnerf_b200", "csrc", "rays.cu")).read()
    m = re.search(r"(struct RayCam \{.*?)\n// hits of this thread's pixel", src, flags=re.S)
    assert m, "rays.cu no longer has the RayCam .. ray_for_pixel section this harness extracts"
    body = m.group(1).replace("#pragma unroll", "")
    os.makedirs(BUILD, exist_ok=True)
    cpp, so = os.path.join(BUILD, "rays_host_emulation.cpp"), os.path.join(BUILD, "rays_host_emulation.so")
    with open(cpp, "w") as f:
        f.write(PRELUDE + body + DRIVER)
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", cpp, "-o", so], check=True)
    return ctypes.CDLL(so)


@pytest.mark.parametrize("name", ["zju", "f64", "f32"])
def test_device_arithmetic_reproduces_the_reference(name):
    lib = _build()
    g = np.load(os.path.join(ROOT, "tests", "golden", f"rays_{name}.npz"))
    H, W, K = int(g["H"]), int(g["W"]), g["K"]
    kinv = np.ascontiguousarray(np.linalg.inv(K).astype(np.float64))
    R, T = np.ascontiguousarray(g["R"], np.float64), np.ascontiguousarray(g["T"], np.float64)
    lo, hi = g["bbox_min"].astype(np.float64), g["bbox_max"].astype(np.float64)
    rays = np.zeros((H * W, 8), np.float32)
    mask = np.zeros(H * W, np.uint8)
    dp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    mode = 2 if (K.dtype == np.float32 and g["R"].dtype == np.float32 and g["T"].dtype == np.float32) else int(K.dtype == np.float32)
    lib.emulate(dp(kinv), mode, dp(R), dp(T), dp(lo), dp(hi), H, W, dp(rays), dp(mask))
    hit = mask.astype(bool)
    assert np.array_equal(hit, g["ray_mask"])
    assert np.array_equal(rays[hit, 0:3], g["rays_o"]) and np.array_equal(rays[hit, 3:6], g["rays_d"])
    assert np.array_equal(rays[hit, 6], g["near"]) and np.array_equal(rays[hit, 7], g["far"])


@pytest.mark.parametrize("H,W,yaw,k_dtype", [(200, 300, 0.0, np.float32), (256, 256, 0.7, np.float64), (97, 131, 2.1, np.float32)])
def test_device_arithmetic_against_oracle(H, W, yaw, k_dtype):
    """Same harness against oracle/rays_oracle.py on the synthetic look-at camera (rotated views, non-square frames)."""
    from occnerf_b200 import synthetic as S
    from oracle import rays_oracle as RO
    lib = _build()
    K, R, T = S.lookat_camera(max(H, W), yaw=yaw)
    K = K.astype(k_dtype)
    K[0, 2], K[1, 2] = W / 2.0, H / 2.0
    R, T = np.ascontiguousarray(R, np.float64), np.ascontiguousarray(T, np.float64)
    bmin, bmax = np.array([-0.95, -1.35, -0.45], np.float32), np.array([0.95, 0.65, 0.45], np.float32)
    want, wmask, _ = RO.frame_rays(H, W, K, R, T, bmin, bmax)
    kinv = np.ascontiguousarray(np.linalg.inv(K).astype(np.float64))
    lo, hi = bmin.astype(np.float64), bmax.astype(np.float64)
    rays, mask = np.zeros((H * W, 8), np.float32), np.zeros(H * W, np.uint8)
    dp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.emulate(dp(kinv), int(K.dtype == np.float32), dp(R), dp(T), dp(lo), dp(hi), H, W, dp(rays), dp(mask))
    hit = mask.astype(bool)
    assert 0.05 * H * W < hit.sum() < H * W
    assert np.array_equal(hit, wmask) and np.array_equal(rays[hit], want)
