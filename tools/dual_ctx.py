"""Experiment: K contexts on ONE GPU, each rendering 1/K of the sample range on its own stream (development aid).
Tests whether the issue-bound trace kernel and the latency-bound shade kernel of different contexts overlap.
usage: dual_ctx.py K spp [key=value ...]"""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from bench import load_scene
K = int(sys.argv[1]); spp = int(sys.argv[2])
opts = [kv.split("=") for kv in sys.argv[3:]]
tris, mats, nodes = load_scene()
ctxs = []
for k in range(K):
    m = yb.CUDAManager().setup(0)
    for key, v in opts:
        m.setOption(key, float(v))
    r = yb.RendererCore(m, 1024, 1024)
    assert m.createRenderProgram("udpt.cl", compiler_opts="-DMIS")
    sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = tris, mats, nodes
    assert r.setup(sc)
    r.enqueueKernels(2)
    ctxs.append((m, r))
bar = threading.Barrier(K + 1)
stats = [None] * K
def work(k):
    m, r = ctxs[k]
    r.frame = k * (spp // K) if hasattr(r, "frame") else 0
    bar.wait()
    stats[k] = r.enqueueKernels(spp // K, reset=True)
    bar.wait()
th = [threading.Thread(target=work, args=(k,)) for k in range(K)]
for t in th: t.start()
bar.wait(); t0 = time.perf_counter(); bar.wait(); t1 = time.perf_counter()
for t in th: t.join()
tot = sum(s.samples for s in stats)
print("K", K, "spp", spp, dict(opts), "wall ms %.1f" % ((t1 - t0) * 1e3), "Msamples/s %.1f" % (tot / (t1 - t0) / 1e6),
      "per-ctx ms", ["%.0f" % s.render_ms for s in stats], "iters", [s.iterations for s in stats])
