"""ctypes binding of the native library (include/yune_cuda.h + include/yune_host.h).

The library is the product; this module only declares its C ABI.  Loading fails loudly when the library has
not been built -- there is no Python or CPU implementation to fall back to.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyune_b200.so")

c_void_pp = C.POINTER(C.c_void_p)


class Stats(C.Structure):
    _fields_ = [("render_ms", C.c_double), ("trace_ms", C.c_double), ("shade_ms", C.c_double), ("tonemap_ms", C.c_double),
                ("samples", C.c_uint64), ("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("box_tests", C.c_uint64), ("tri_tests", C.c_uint64),
                ("iterations", C.c_uint32), ("kernel_launches", C.c_uint32), ("trace_launches", C.c_uint32), ("timed_iterations", C.c_uint32),
                ("diffuse_visits", C.c_uint64), ("specular_visits", C.c_uint64), ("regenerations", C.c_uint64), ("slot_visits", C.c_uint64),
                ("steady_iterations", C.c_uint32), ("steady_timed_iterations", C.c_uint32), ("steady_extend_rays", C.c_uint64),
                ("steady_shadow_rays", C.c_uint64), ("steady_trace_ms", C.c_double), ("steady_shade_ms", C.c_double),
                ("carried_paths", C.c_uint32), ("reserved0", C.c_uint32), ("finish_ms", C.c_double), ("sort_ms", C.c_double)]


class GroupStats(C.Structure):
    _fields_ = [("n_devices", C.c_int), ("render_ms_max", C.c_double), ("render_ms_min", C.c_double), ("reduce_ms", C.c_double),
                ("samples", C.c_uint64), ("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64)]


# name -> (restype, argtypes); every symbol include/*.h declares is listed here (tests check the export table against it)
CUDA_API = {
    "yune_setup": (C.c_int, [C.c_int, c_void_pp]),
    "yune_destroy": (None, [C.c_void_p]),
    "yune_last_error": (C.c_char_p, [C.c_void_p]),
    "yune_create_render_program": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "yune_create_postproc_program": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "yune_setup_vertex_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yune_setup_mat_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yune_setup_bvh_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yune_build_bvh_on_device": (C.c_int, [C.c_void_p, C.c_int]),
    "yune_bvh_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "yune_read_bvh_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yune_setup_camera_buffer": (C.c_int, [C.c_void_p, C.c_void_p]),
    "yune_setup_image_buffers": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "yune_set_light_sources": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yune_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "yune_get_option": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]),
    "yune_render": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int]),
    "yune_finish": (C.c_int, [C.c_void_p]),
    "yune_tonemap": (C.c_int, [C.c_void_p]),
    "yune_read_hdr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "yune_read_sum": (C.c_int, [C.c_void_p, C.c_void_p]),
    "yune_read_ldr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "yune_write_sum": (C.c_int, [C.c_void_p, C.c_void_p]),
    "yune_sum_device_ptr": (C.c_int, [C.c_void_p, c_void_pp, C.POINTER(C.c_size_t)]),
    "yune_read_sum_fixed": (C.c_int, [C.c_void_p, C.c_void_p]),
    "yune_sum_fixed_device_ptr": (C.c_int, [C.c_void_p, c_void_pp, C.POINTER(C.c_size_t)]),
    "yune_sum_refresh": (C.c_int, [C.c_void_p]),
    "yune_stream": (C.c_int, [C.c_void_p, c_void_pp]),
    "yune_synchronize": (C.c_int, [C.c_void_p]),
    "yune_trace_primary": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "yune_trace_rays": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "yune_debug_capture_rays": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "yune_debug_read_captured": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "yune_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "yune_shard_samples": (None, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "yune_group_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), c_void_pp]),
    "yune_group_destroy": (None, [C.c_void_p]),
    "yune_group_size": (C.c_int, [C.c_void_p]),
    "yune_group_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "yune_group_last_error": (C.c_char_p, [C.c_void_p]),
    "yune_group_create_render_program": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "yune_group_create_postproc_program": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "yune_group_setup_vertex_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yune_group_setup_mat_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yune_group_setup_bvh_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yune_group_build_bvh_on_device": (C.c_int, [C.c_void_p, C.c_int]),
    "yune_group_setup_camera_buffer": (C.c_int, [C.c_void_p, C.c_void_p]),
    "yune_group_setup_image_buffers": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "yune_group_set_light_sources": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yune_group_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "yune_group_render": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int]),
    "yune_group_reduce": (C.c_int, [C.c_void_p, C.c_int]),
    "yune_group_get_stats": (C.c_int, [C.c_void_p, C.POINTER(GroupStats)]),
}
HOST_API = {
    "yune_scene_create": (C.c_void_p, []),
    "yune_scene_destroy": (None, [C.c_void_p]),
    "yune_scene_last_error": (C.c_char_p, [C.c_void_p]),
    "yune_scene_load_model": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "yune_scene_set_geometry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "yune_scene_load_bvh": (C.c_int, [C.c_void_p, C.c_int]),
    "yune_scene_reload_mat_file": (C.c_int, [C.c_void_p]),
    "yune_scene_num_triangles": (C.c_int, [C.c_void_p]),
    "yune_scene_num_materials": (C.c_int, [C.c_void_p]),
    "yune_scene_num_bvh_nodes": (C.c_int, [C.c_void_p]),
    "yune_scene_vert_data": (C.c_void_p, [C.c_void_p]),
    "yune_scene_mat_data": (C.c_void_p, [C.c_void_p]),
    "yune_scene_mat_data_mut": (C.c_void_p, [C.c_void_p]),
    "yune_scene_bvh_data": (C.c_void_p, [C.c_void_p]),
    "yune_scene_root_aabb": (None, [C.c_void_p, C.c_void_p]),
    "yune_camera_default": (None, [C.c_float, C.c_void_p]),
    "yune_camera_set": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "yune_camera_create": (C.c_void_p, [C.c_float]),
    "yune_camera_destroy": (None, [C.c_void_p]),
    "yune_camera_set_orientation": (None, [C.c_void_p, C.c_void_p, C.c_float, C.c_float]),
    "yune_camera_reset": (None, [C.c_void_p]),
    "yune_camera_is_changed": (C.c_int, [C.c_void_p]),
    "yune_camera_set_buffer": (None, [C.c_void_p, C.c_void_p]),
    "yune_write_image": (C.c_int, [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int]),
}

_lib = None


def load():
    """Return the loaded library (built on first use if sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("YUNE_B200_LIB", LIB_PATH)        # development: A/B a differently-compiled build of the SAME sources
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run `python -m yune_b200.build` (nvcc, sm_100a). "
                           "There is no CPU implementation of the render path." % path)
    lib = C.CDLL(path)
    for table in (CUDA_API, HOST_API):
        for name, (res, args) in table.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib
