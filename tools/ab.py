"""A/B differently-compiled builds of the library on the bench workload (development aid; run under gpurun).

    python tools/ab.py [--spp 256] [--size 1024] [--opt k=v ...] build/variants/a.so build/variants/b.so ...

Every library runs in its own process (YUNE_B200_LIB): C2 at the given spp, stage timing on every 4th iteration.  Prints one
JSON line per library: Msamples/s, average trace / shade launch, iterations, and the image mean (a sanity check, not parity)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("YUNE_AB_CHILD"):
    sys.path.insert(0, ROOT)
    import numpy as np
    import yune_b200 as yb
    from bench import build_scene, CONFIGS
    spp, size = int(sys.argv[1]), int(sys.argv[2])
    opts = dict(kv.split("=") for kv in sys.argv[3:])
    cfg = CONFIGS[opts.pop("config", "c2")]
    tris, mats, nodes, lights = build_scene(cfg)
    m = yb.CUDAManager().setup(0)
    for k, v in opts.items():
        m.setOption(k, float(v))
    W, H = (size, size) if cfg["W"] == cfg["H"] else (size, size * cfg["H"] // cfg["W"])
    r = yb.RendererCore(m, W, H)
    assert m.createRenderProgram(cfg["kernel"], compiler_opts=cfg["opts"])
    sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = tris, mats, nodes
    assert r.setup(sc), m.last_message
    if lights is not None:
        assert m.setLightSources(lights)
    m.setOption("oren_nayar", 1 if cfg.get("oren_nayar") else 0)
    r.enqueueKernels(8)
    m.setOption("time_stages", 4)
    best = None
    for rep in range(2):
        st = r.enqueueKernels(spp, reset=True)
        n = max(st.timed_iterations, 1)
        res = dict(msamples_s=round(st.samples / st.render_ms / 1e3, 1), ms=round(st.render_ms, 2), trace_ms=round(st.trace_ms / n, 4), shade_ms=round(st.shade_ms / n, 4), sort_ms=round(st.sort_ms / n, 4),
                   iterations=int(st.iterations))
        if best is None or res["msamples_s"] > best["msamples_s"]:
            best = res
    best["mean"] = float(r.readHDR()[..., :3].mean())
    print(json.dumps(best))
    sys.exit(0)
args = sys.argv[1:]
spp, size, opts, libs = 256, 1024, [], []
while args:
    a = args.pop(0)
    if a == "--spp": spp = int(args.pop(0))
    elif a == "--size": size = int(args.pop(0))
    elif a == "--opt": opts.append(args.pop(0))
    else: libs.append(a)
for lib in libs:
    env = dict(os.environ, YUNE_AB_CHILD="1")
    if lib != "default":
        env["YUNE_B200_LIB"] = os.path.abspath(lib)
    p = subprocess.run([sys.executable, os.path.abspath(__file__), str(spp), str(size)] + opts, env=env, capture_output=True, text=True, timeout=600)
    line = p.stdout.strip().splitlines()[-1] if p.returncode == 0 and p.stdout.strip() else "FAILED rc=%d %s" % (p.returncode, p.stderr[-400:])
    print("%-28s %s" % (os.path.basename(lib), line), flush=True)
