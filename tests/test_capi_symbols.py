"""The drop-in boundary: the library loads, exports every entry point include/*.h declares, and refuses to run without
a B200 instead of falling back to the CPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from yune_b200 import _native
from tests.helpers import ROOT


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yune_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    names = _declared("yune_cuda.h") + _declared("yune_host.h")
    assert len(names) > 40
    for n in names:
        assert hasattr(lib, n), "include/ declares %s but the library does not export it" % n
    assert set(names) == set(_native.CUDA_API) | set(_native.HOST_API), "ctypes tables and headers disagree"


def test_library_is_sm100a_native_code():
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_pod_sizes_match_reference_layouts():
    from yune_b200 import api
    assert (api.TRI_DTYPE.itemsize, api.NODE_DTYPE.itemsize, api.MAT_DTYPE.itemsize, api.CAM_DTYPE.itemsize, api.QUAD_DTYPE.itemsize) == (112, 80, 80, 80, 128)
    assert api.TRI_DTYPE.fields["matID"][1] == 96 and api.NODE_DTYPE.fields["child_idx"][1] == 72 and api.MAT_DTYPE.fields["is_specular"][1] == 72


def test_no_cpu_fallback():
    """Without a GPU yune_setup must fail with YUNE_ERR_NODEVICE; with one it must succeed -- never a silent CPU path."""
    import torch
    lib = _native.load()
    ctx = C.c_void_p()
    rc = lib.yune_setup(0, C.byref(ctx))
    if torch.cuda.is_available():
        assert rc == 0
        lib.yune_destroy(ctx)
    else:
        assert rc == -5 and b"no CPU fallback" in lib.yune_last_error(None)
        import yune_b200 as yb
        with pytest.raises(yb.YuneError):
            yb.CUDAManager().setup(0)
        # the C++ front end says so too and writes nothing
        import subprocess
        from tests.helpers import ROOT
        out = os.path.join(ROOT, "gpurun_out", "_never_written.png")
        p = subprocess.run([os.path.join(ROOT, "yune_b200", "yune_headless"), "--obj", "missing.obj", "--spp", "1", "--out", out], capture_output=True, text=True)
        assert p.returncode == 1 and "no CPU fallback" in p.stderr and not os.path.exists(out)


def test_argument_errors_do_not_crash():
    lib = _native.load()
    assert lib.yune_setup(0, None) == -1
    assert lib.yune_render(None, 0, 1, 1, 0, 1) == -1
    assert lib.yune_get_stats(None, None) == -1
    # round-2 entry points
    assert lib.yune_finish(None) == -1
    assert lib.yune_build_bvh_on_device(None, 2) == -1
    assert lib.yune_bvh_info(None, None, None, None, None) == -1
    assert lib.yune_read_bvh_buffer(None, None, 0) == -1


def test_group_api_without_a_device_and_shard_rule():
    """yune_group_* (SURVEY.md 8b): argument errors and the no-device answer do not need a GPU; the split rule is the one the
    torch.distributed path uses (yune_b200/dist.py) -- contiguous, balanced, disjoint, covering."""
    import torch
    from yune_b200.dist import shard_samples as py_shard
    import yune_b200 as yb
    lib = _native.load()
    assert lib.yune_group_create(2, None, None) == -1
    assert lib.yune_group_size(None) == 0 and lib.yune_group_ctx(None, 0) is None
    assert lib.yune_group_render(None, 0, 1, 1, 0, 1) == -1 and lib.yune_group_reduce(None, 0) == -1
    if not torch.cuda.is_available():
        g = C.c_void_p()
        assert lib.yune_group_create(2, None, C.byref(g)) == -5 and b"no CPU fallback" in lib.yune_group_last_error(None)
        with pytest.raises(yb.YuneError):
            yb.CUDAGroup(2)
    for total, begin in ((16384, 0), (1024, 512), (7, 3), (0, 0), (5, 0)):
        for n in (1, 2, 3, 4, 8):
            got = [yb.shard_samples(begin, total, r, n) for r in range(n)]
            want = [py_shard(total, r, n) for r in range(n)]
            assert [(b - begin, c) for b, c in got] == want
            assert sum(c for _, c in got) == total and got[0][0] == begin
            assert all(got[r][0] + got[r][1] == got[r + 1][0] for r in range(n - 1))
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
