#!/usr/bin/env python3
"""tests/golden/make_golden.py -- regenerates the committed fixtures from the REFERENCE itself.

Runs only where /root/reference and oracle/_ref exist (the development container).  Everything it writes is
data produced by running the reference's own code (its host sources compiled by oracle/ref_host_api.cpp,
its OpenCL kernels compiled as C++ by oracle/gen_ref_kernels.py); the GPU box, which has no /root/reference,
only ever reads the fixtures.

  scene_<name>.npz   vert_data (112 B records), mat_data (80 B), bvh (80 B) exactly as the reference builds them,
                     with the bytes the reference leaves uninitialised normalised (TriangleGPU.pad = 0,
                     vert_list slots >= vert_len = -1) so the fixture is deterministic.
  primary_<name>.npz primary-ray hit ids of createRay+traceRay (udpt.cl:213-286): pixel centres and the
                     kernel's own jitter for rand = 12345, default camera.
  kat.npz            known-answer vectors of small functions (wang_hash, xor_shift, tonemap ramp).
  hdr_<cfg>.npz      reference HDR frames (running mean) for fixed per-frame `rand` lists, small resolutions.
"""
import ctypes as C
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.refbind import RefHost, RefKernels, default_cam_array, frame_rands  # noqa: E402

GEOM = "/root/reference/geometry"


def normalise(tris, nodes):
    tris = tris.copy(); nodes = nodes.copy()
    tris["pad"] = 0
    for j in range(10):
        nodes["vert_list"][:, j] = np.where(j < np.maximum(nodes["vert_len"], 0), nodes["vert_list"][:, j], -1)
    return tris, nodes


def teapot_transmissive(tris, mats):
    """Config C2: the teapot's triangles (material 'teapot', index 3) use 'mirror-transmissive' (index 4).
    Equivalent to editing 'usemtl teapot' at geometry/cornellbox-teapot.obj:6670; geometry and BVH are unchanged."""
    t = tris.copy()
    t["matID"] = np.where(t["matID"] == 3, 4, t["matID"])
    return t


def main():
    host, kern = RefHost(), RefKernels()
    cam = default_cam_array()
    out = {}
    for name, fn in (("cornellbox", "cornellbox.obj"), ("teapot", "cornellbox-teapot.obj")):
        tris, mats, nodes, root = host.load(os.path.join(GEOM, fn))
        tris, nodes = normalise(tris, nodes)
        np.savez_compressed(os.path.join(HERE, "scene_%s.npz" % name), vert_data=tris, mat_data=mats, bvh=nodes, root=root)
        out[name] = (tris, mats, nodes)
        W = 256
        prim = {}
        for jm in (0, 1):
            tri, light, t, od = kern.primary("udpt", cam, tris, nodes, 12345, jm, W, W)
            prim["tri_j%d" % jm] = tri.astype(np.int16 if tris.size < 32767 else np.int32)
            prim["light_j%d" % jm] = light.astype(np.int8)
            prim["t_j%d" % jm] = t
        np.savez_compressed(os.path.join(HERE, "primary_%s.npz" % name), width=W, height=W, rand=12345, **prim)
        print(name, tris.size, "tris", nodes.size, "nodes")

    # known answers
    seeds = np.array([0, 1, 2, 61, 12345, 0xdeadbeef, 0xffffffff, 0x80000000], np.uint32)
    wh = np.array([kern.lib.yref_udpt_wang_hash(C.c_uint(int(s))) & 0xffffffff for s in seeds], np.uint32)
    xs = np.array([kern.lib.yref_udpt_xor_shift(C.c_uint(int(s))) & 0xffffffff for s in seeds], np.uint32)
    ramp = np.zeros((4, 64, 4), np.float32)
    ramp[..., 0] = np.linspace(0, 4, 64)[None, :]
    ramp[..., 1] = np.linspace(0, 2, 64)[None, :] ** 2
    ramp[..., 2] = np.linspace(0.001, 30, 64)[None, :]
    ramp[..., 3] = np.arange(1, 5, dtype=np.float32)[:, None]
    tm = kern.tonemap(ramp)
    np.savez_compressed(os.path.join(HERE, "kat.npz"), seeds=seeds, wang_hash=wh, xor_shift=xs, tonemap_in=ramp, tonemap_out=tm)

    # reference HDR frames (running mean, count in alpha), rand list = mt19937(12345) stand-in for the clock seed
    tris, mats, nodes = out["cornellbox"]
    for cfg, variant, W, spp in (("c1_udpt_128", "udpt", 128, 64), ("c1_udptmis_128", "udpt_mis", 128, 64)):
        img = kern.render(variant, cam, tris, mats, nodes, W, W, frame_rands(12345, spp))
        np.savez_compressed(os.path.join(HERE, "hdr_%s.npz" % cfg), image=img, spp=spp, seed=12345)
        print(cfg, "mean", img[..., :3][np.isfinite(img[..., :3])].mean())
    tris, mats, nodes = out["teapot"]
    tt = teapot_transmissive(tris, mats)
    for cfg, variant, tr, W, spp in (("c2_udptmis_96", "udpt_mis", tt, 96, 64), ("c3_bdpt_64", "bdpt", tris, 64, 32)):
        img = kern.render(variant, cam, tr, mats, nodes, W, W, frame_rands(12345, spp))
        np.savez_compressed(os.path.join(HERE, "hdr_%s.npz" % cfg), image=img, spp=spp, seed=12345)
        print(cfg, "mean", img[..., :3][np.isfinite(img[..., :3])].mean())


if __name__ == "__main__":
    main()
