"""configs[4] ("C5"): Cornell-teapot 3840x2160, a total spp budget PARTITIONED by sample index over the ranks (strong scaling),
one NCCL SUM-reduce of the 132.7 MB fp32 accumulation buffer to rank 0, divide + tonemap there.  Measurement aid:

    python tools/run_c5.py [total_spp]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_c5.py [total_spp]

Prints one JSON line on rank 0 (device time: max over ranks of render, + reduce + tonemap on rank 0)."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import yune_b200 as yb
from yune_b200.dist import shard_samples, reduce_sum_to_root, sum_buffer_as_tensor
from bench import load_scene

W, H = 3840, 2160
total_spp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tris, mats, nodes = load_scene()
m = yb.CUDAManager().setup(local)
r = yb.RendererCore(m, W, H); r.seed = 12345
assert m.createRenderProgram("udpt.cl", compiler_opts="-DMIS") and m.createPostProcProgram("tonemap.cl")
sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = tris, mats, nodes
assert r.setup(sc)
begin, count = shard_samples(total_spp, rank, world)
sum_t = sum_buffer_as_tensor(r) if world > 1 else None
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = []
for rep in range(3):                                  # first repetition = warm-up (pool allocation, NCCL channels)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    m.check(r._lib.yune_render(r._ctx, begin, count, 1, r.seed, 1))
    st = yb._native.Stats(); r._lib.yune_get_stats(r._ctx, C.byref(st))
    ms = torch.tensor([st.render_ms], device="cuda")
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    red = 0.0
    if world > 1:
        ev0.record(); reduce_sum_to_root(sum_t); ev1.record(); torch.cuda.synchronize(); red = ev0.elapsed_time(ev1)
    tm = 0.0
    if rank == 0:
        m.check(r._lib.yune_tonemap(r._ctx))
        st2 = yb._native.Stats(); r._lib.yune_get_stats(r._ctx, C.byref(st2)); tm = st2.tonemap_ms
    res.append((float(ms.item()), red, tm))
if rank == 0:
    s = r.readSum()
    render_ms, red, tm = res[-1]
    total = render_ms + red + tm
    print(json.dumps({"config": "C5 (reduced spp)", "width": W, "height": H, "total_spp": total_spp, "n_gpus": world, "spp_per_rank": count,
                      "render_ms_max_over_ranks": render_ms, "reduce_ms": red, "reduce_bytes": W * H * 16, "tonemap_ms": tm,
                      "msamples_s": W * H * total_spp / total / 1e3, "samples_per_pixel_in_sum": [float(s[..., 3].min()), float(s[..., 3].max())],
                      "nonfinite_pixels": int((~np.isfinite(s)).sum())}))
if world > 1:
    dist.destroy_process_group()
