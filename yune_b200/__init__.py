"""yune_b200 -- B200-native path-tracing hot path behind Yune's host surface.

Python is harness only: the product is yune_b200/libyune_b200.so (CUDA kernels + C ABI, include/*.h).
`Scene`, `CUDAManager` and `RendererCore` mirror the reference classes of the same roles
(include/Scene.h, include/CLManager.h, include/RendererCore.h) on top of that C ABI.
"""
from .api import Scene, Camera, CUDAManager, CUDAGroup, RendererCore, shard_samples, YuneError, default_camera, quad_light, write_image, LIGHT_UDPT, LIGHT_BDPT  # noqa: F401
