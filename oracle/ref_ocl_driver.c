/* ref_ocl_driver.c -- TEST / MEASUREMENT INFRASTRUCTURE, not part of the product.
 *
 * Runs the reference's OWN kernel text (kernels/legacy/udpt.cl, unmodified apart from the '#yune-preproc' header lines that
 * the reference's loader strips too, src/CLManager.cpp:182-204) on an OpenCL device of the box -- on a B200 that is NVIDIA's
 * OpenCL driver -- with the reference's launch protocol (src/RendererCore.cpp:248-306, 508-606: 14 kernel arguments, one
 * NDRange over the image per frame, a fresh `rand` per frame, two images that swap roles).  SURVEY.md 8c / 8d call this the
 * opportunistic second baseline: the reference on the same GPU.
 *
 * The kernel text is embedded at build time from the reference tree where it lies (oracle/gen_ref_ocl.py writes the string
 * literal into a temporary directory; the binary goes to oracle/_ref/, which is git-ignored).  No reference source is copied
 * into the repository.
 *
 * There are no OpenCL headers in the image, so the few types / constants used are declared here.  The OpenCL library is bound
 * at run time: the ICD loader (libOpenCL.so.1) with OCL_ICD_FILENAMES / OCL_ICD_VENDORS pointing at the vendor library when
 * /etc/OpenCL/vendors is absent, else the vendor library itself (libnvidia-opencl.so.1) through its ICD dispatch table.
 *
 *   yune_ref_ocl SCENE_DIR W H FRAMES GI [build options]      SCENE_DIR holds tris.bin mats.bin nodes.bin cam.bin rands.bin
 * prints one JSON line; writes SCENE_DIR/image.bin (W*H float4, the running mean + count, what the reference displays).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int32_t cl_int; typedef uint32_t cl_uint; typedef uint64_t cl_ulong; typedef cl_ulong cl_bitfield;
typedef struct _cl_platform_id* cl_platform_id; typedef struct _cl_device_id* cl_device_id; typedef struct _cl_context* cl_context;
typedef struct _cl_command_queue* cl_command_queue; typedef struct _cl_mem* cl_mem; typedef struct _cl_program* cl_program;
typedef struct _cl_kernel* cl_kernel; typedef struct _cl_event* cl_event;
typedef struct { cl_uint image_channel_order, image_channel_data_type; } cl_image_format;
#define CL_DEVICE_TYPE_GPU (1 << 2)
#define CL_DEVICE_TYPE_ALL 0xFFFFFFFF
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_COPY_HOST_PTR (1 << 5)
#define CL_RGBA 0x10B5
#define CL_FLOAT 0x10DE
#define CL_PLATFORM_NAME 0x0902
#define CL_DEVICE_NAME 0x102B
#define CL_DEVICE_VERSION 0x102F
#define CL_PROGRAM_BUILD_LOG 0x1183
#define CL_QUEUE_PROFILING_ENABLE (1 << 1)
#define CL_PROFILING_COMMAND_START 0x1282
#define CL_PROFILING_COMMAND_END 0x1283

/* entry points, in the order of the ICD dispatch table (cl_icd.h, OpenCL 1.0 block) */
enum { I_GetPlatformIDs = 0, I_GetPlatformInfo = 1, I_GetDeviceIDs = 2, I_GetDeviceInfo = 3, I_CreateContext = 4, I_CreateCommandQueue = 9,
       I_CreateBuffer = 14, I_CreateImage2D = 15, I_ReleaseMemObject = 18, I_CreateProgramWithSource = 26, I_BuildProgram = 30,
       I_GetProgramBuildInfo = 33, I_CreateKernel = 34, I_SetKernelArg = 38, I_ReleaseEvent = 44, I_GetEventProfilingInfo = 45, I_Finish = 47,
       I_EnqueueReadImage = 51, I_EnqueueWriteImage = 52, I_EnqueueNDRangeKernel = 59, I_COUNT = 60 };
static const char* k_names[I_COUNT];
static void* fn[I_COUNT];
typedef cl_int (*F_GetPlatformIDs)(cl_uint, cl_platform_id*, cl_uint*);
typedef cl_int (*F_GetPlatformInfo)(cl_platform_id, cl_uint, size_t, void*, size_t*);
typedef cl_int (*F_GetDeviceIDs)(cl_platform_id, cl_bitfield, cl_uint, cl_device_id*, cl_uint*);
typedef cl_int (*F_GetDeviceInfo)(cl_device_id, cl_uint, size_t, void*, size_t*);
typedef cl_context (*F_CreateContext)(const intptr_t*, cl_uint, const cl_device_id*, void*, void*, cl_int*);
typedef cl_command_queue (*F_CreateCommandQueue)(cl_context, cl_device_id, cl_bitfield, cl_int*);
typedef cl_mem (*F_CreateBuffer)(cl_context, cl_bitfield, size_t, void*, cl_int*);
typedef cl_mem (*F_CreateImage2D)(cl_context, cl_bitfield, const cl_image_format*, size_t, size_t, size_t, void*, cl_int*);
typedef cl_program (*F_CreateProgramWithSource)(cl_context, cl_uint, const char**, const size_t*, cl_int*);
typedef cl_int (*F_BuildProgram)(cl_program, cl_uint, const cl_device_id*, const char*, void*, void*);
typedef cl_int (*F_GetProgramBuildInfo)(cl_program, cl_device_id, cl_uint, size_t, void*, size_t*);
typedef cl_kernel (*F_CreateKernel)(cl_program, const char*, cl_int*);
typedef cl_int (*F_SetKernelArg)(cl_kernel, cl_uint, size_t, const void*);
typedef cl_int (*F_ReleaseEvent)(cl_event);
typedef cl_int (*F_GetEventProfilingInfo)(cl_event, cl_uint, size_t, void*, size_t*);
typedef cl_int (*F_Finish)(cl_command_queue);
typedef cl_int (*F_EnqueueReadImage)(cl_command_queue, cl_mem, cl_uint, const size_t*, const size_t*, size_t, size_t, void*, cl_uint, const cl_event*, cl_event*);
typedef cl_int (*F_EnqueueWriteImage)(cl_command_queue, cl_mem, cl_uint, const size_t*, const size_t*, size_t, size_t, const void*, cl_uint, const cl_event*, cl_event*);
typedef cl_int (*F_EnqueueNDRangeKernel)(cl_command_queue, cl_kernel, cl_uint, const size_t*, const size_t*, const size_t*, cl_uint, const cl_event*, cl_event*);
#define CL(name) ((F_##name)fn[I_##name])

extern const char* yune_ref_kernel_text;      /* generated at build time from the reference tree */

static void fail(const char* what, long code)
{
    printf("{\"impl\": \"reference-opencl\", \"unavailable\": \"%s (%ld)\"}\n", what, code);
    exit(0);
}

static void* read_file(const char* dir, const char* name, size_t* n)
{
    char p[1024]; snprintf(p, sizeof p, "%s/%s", dir, name);
    FILE* f = fopen(p, "rb"); if (!f) fail("cannot open scene file", 0);
    fseek(f, 0, SEEK_END); *n = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    void* b = malloc(*n ? *n : 1);
    if (fread(b, 1, *n, f) != *n) fail("short read", 0);
    fclose(f);
    return b;
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* Bind the entry points.  Returns a platform that has a GPU device, or NULL. */
static cl_platform_id bind_opencl(char* how, size_t how_n)
{
    k_names[I_GetPlatformIDs] = "clGetPlatformIDs"; k_names[I_GetPlatformInfo] = "clGetPlatformInfo"; k_names[I_GetDeviceIDs] = "clGetDeviceIDs";
    k_names[I_GetDeviceInfo] = "clGetDeviceInfo"; k_names[I_CreateContext] = "clCreateContext"; k_names[I_CreateCommandQueue] = "clCreateCommandQueue";
    k_names[I_CreateBuffer] = "clCreateBuffer"; k_names[I_CreateImage2D] = "clCreateImage2D"; k_names[I_ReleaseMemObject] = "clReleaseMemObject";
    k_names[I_CreateProgramWithSource] = "clCreateProgramWithSource"; k_names[I_BuildProgram] = "clBuildProgram";
    k_names[I_GetProgramBuildInfo] = "clGetProgramBuildInfo"; k_names[I_CreateKernel] = "clCreateKernel"; k_names[I_SetKernelArg] = "clSetKernelArg";
    k_names[I_ReleaseEvent] = "clReleaseEvent"; k_names[I_GetEventProfilingInfo] = "clGetEventProfilingInfo"; k_names[I_Finish] = "clFinish";
    k_names[I_EnqueueReadImage] = "clEnqueueReadImage"; k_names[I_EnqueueWriteImage] = "clEnqueueWriteImage"; k_names[I_EnqueueNDRangeKernel] = "clEnqueueNDRangeKernel";
    const char* vendor_libs[] = {"libnvidia-opencl.so.1", "/usr/lib/libnvidia-opencl.so.1", "/usr/local/nvidia/lib/libnvidia-opencl.so.1", "/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so.1", NULL};
    /* 1. the ICD loader; point it at the vendor library when the registry directory is missing */
    const char* loaders[] = {"libOpenCL.so.1", "/usr/local/cuda/lib64/libOpenCL.so.1", "/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL.so.1", NULL};
    for (int v = 0; vendor_libs[v]; v++) {
        void* probe = dlopen(vendor_libs[v], RTLD_NOW | RTLD_GLOBAL);
        if (!probe) continue;
        setenv("OCL_ICD_FILENAMES", vendor_libs[v], 0);
        for (int l = 0; loaders[l]; l++) {
            void* h = dlopen(loaders[l], RTLD_NOW);
            if (!h) continue;
            int ok = 1;
            for (int i = 0; i < I_COUNT; i++) if (k_names[i]) { fn[i] = dlsym(h, k_names[i]); if (!fn[i]) ok = 0; }
            if (!ok) continue;
            cl_platform_id plats[8]; cl_uint np = 0;
            if (CL(GetPlatformIDs)(8, plats, &np) == 0)
                for (cl_uint p = 0; p < np && p < 8; p++) {
                    cl_device_id d; cl_uint nd = 0;
                    if (CL(GetDeviceIDs)(plats[p], CL_DEVICE_TYPE_GPU, 1, &d, &nd) == 0 && nd > 0) { snprintf(how, how_n, "ICD loader %s -> %s", loaders[l], vendor_libs[v]); return plats[p]; }
                }
        }
        /* 2. the vendor library itself: platforms from clIcdGetPlatformIDsKHR, every other entry point from the dispatch table
         *    that each OpenCL object starts with */
        typedef void* (*F_GetExt)(const char*);
        F_GetExt get_ext = (F_GetExt)dlsym(probe, "clGetExtensionFunctionAddress");
        if (!get_ext) continue;
        F_GetPlatformIDs icd_platforms = (F_GetPlatformIDs)get_ext("clIcdGetPlatformIDsKHR");
        if (!icd_platforms) continue;
        cl_platform_id plats[8]; cl_uint np = 0;
        if (icd_platforms(8, plats, &np) != 0 || np == 0) continue;
        void** table = *(void***)plats[0];
        for (int i = 0; i < I_COUNT; i++) fn[i] = table[i];
        cl_device_id d; cl_uint nd = 0;
        if (CL(GetDeviceIDs)(plats[0], CL_DEVICE_TYPE_GPU, 1, &d, &nd) == 0 && nd > 0) { snprintf(how, how_n, "dispatch table of %s", vendor_libs[v]); return plats[0]; }
    }
    return NULL;
}

int main(int argc, char** argv)
{
    if (argc < 6) { fprintf(stderr, "usage: yune_ref_ocl SCENE_DIR W H FRAMES GI [build options]\n"); return 2; }
    const char* dir = argv[1];
    const int W = atoi(argv[2]), H = atoi(argv[3]), frames = atoi(argv[4]), gi = atoi(argv[5]);
    const char* opts = argc > 6 ? argv[6] : "";
    char how[512] = "";
    cl_platform_id plat = bind_opencl(how, sizeof how);
    if (!plat) fail("no OpenCL platform with a GPU device could be bound", 0);
    cl_device_id dev; cl_uint nd = 0; cl_int err = CL(GetDeviceIDs)(plat, CL_DEVICE_TYPE_GPU, 1, &dev, &nd);
    if (err || !nd) fail("clGetDeviceIDs", err);
    char pname[256] = "", dname[256] = "", dver[256] = "";
    CL(GetPlatformInfo)(plat, CL_PLATFORM_NAME, sizeof pname, pname, NULL);
    CL(GetDeviceInfo)(dev, CL_DEVICE_NAME, sizeof dname, dname, NULL);
    CL(GetDeviceInfo)(dev, CL_DEVICE_VERSION, sizeof dver, dver, NULL);
    cl_context ctx = CL(CreateContext)(NULL, 1, &dev, NULL, NULL, &err); if (err) fail("clCreateContext", err);
    cl_command_queue q = CL(CreateCommandQueue)(ctx, dev, CL_QUEUE_PROFILING_ENABLE, &err); if (err) fail("clCreateCommandQueue", err);

    size_t n_tris_b, n_mats_b, n_nodes_b, n_cam_b, n_rands_b;
    void* tris = read_file(dir, "tris.bin", &n_tris_b); void* mats = read_file(dir, "mats.bin", &n_mats_b);
    void* nodes = read_file(dir, "nodes.bin", &n_nodes_b); void* cam = read_file(dir, "cam.bin", &n_cam_b);
    cl_uint* rands = (cl_uint*)read_file(dir, "rands.bin", &n_rands_b);
    if ((int)(n_rands_b / 4) < frames + 1) fail("rands.bin holds fewer values than frames + 1", 0);
    const cl_int scene_size = (cl_int)(n_tris_b / 112), bvh_size = (cl_int)(n_nodes_b / 80);       /* include/CL_headers.h:67-92 */
    cl_mem b_tris = CL(CreateBuffer)(ctx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, n_tris_b, tris, &err); if (err) fail("vertex buffer", err);
    cl_mem b_mats = CL(CreateBuffer)(ctx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, n_mats_b, mats, &err); if (err) fail("material buffer", err);
    cl_mem b_nodes = CL(CreateBuffer)(ctx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, n_nodes_b, nodes, &err); if (err) fail("bvh buffer", err);
    cl_mem b_cam = CL(CreateBuffer)(ctx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, n_cam_b, cam, &err); if (err) fail("camera buffer", err);
    const cl_image_format fmt = {CL_RGBA, CL_FLOAT};
    cl_mem img[2];
    for (int i = 0; i < 2; i++) { img[i] = CL(CreateImage2D)(ctx, CL_MEM_READ_WRITE, &fmt, (size_t)W, (size_t)H, 0, NULL, &err); if (err) fail("clCreateImage2D", err); }

    const char* src = yune_ref_kernel_text;
    double t0 = now_s();
    cl_program prog = CL(CreateProgramWithSource)(ctx, 1, &src, NULL, &err); if (err) fail("clCreateProgramWithSource", err);
    err = CL(BuildProgram)(prog, 1, &dev, opts[0] ? opts : NULL, NULL, NULL);
    if (err) {
        static char log[1 << 16]; size_t n = 0;
        CL(GetProgramBuildInfo)(prog, dev, CL_PROGRAM_BUILD_LOG, sizeof log - 1, log, &n);
        fprintf(stderr, "build log:\n%.*s\n", (int)n, log);
        fail("clBuildProgram", err);
    }
    const double build_s = now_s() - t0;
    cl_kernel k = CL(CreateKernel)(prog, "pathtracer", &err); if (err) fail("clCreateKernel", err);

    /* the 14 arguments of template/kernel.cl:66-68, set as RendererCore::updateRenderKernelArgs / setup do */
    cl_int one = 1, zero = 0, gi_check = gi;
    CL(SetKernelArg)(k, 2, sizeof(cl_mem), &b_cam); CL(SetKernelArg)(k, 3, sizeof(cl_int), &scene_size); CL(SetKernelArg)(k, 4, sizeof(cl_mem), &b_tris);
    CL(SetKernelArg)(k, 5, sizeof(cl_mem), &b_mats); CL(SetKernelArg)(k, 6, sizeof(cl_int), &bvh_size); CL(SetKernelArg)(k, 7, sizeof(cl_mem), &b_nodes);
    CL(SetKernelArg)(k, 8, sizeof(cl_int), &gi_check); CL(SetKernelArg)(k, 11, sizeof(cl_int), &zero); CL(SetKernelArg)(k, 12, sizeof(cl_int), &one); CL(SetKernelArg)(k, 13, sizeof(cl_int), &one);
    const size_t gws[2] = {(size_t)((W + 15) / 16 * 16), (size_t)((H + 15) / 16 * 16)};
    double kernel_ms = 0.0, wall0 = 0.0;
    int cur = 0;
    /* frame -1 is an untimed warm-up (its image is overwritten: frame 0 runs with reset = 1) */
    for (int f = -1; f < frames; f++) {
        if (f == 0) { CL(Finish)(q); wall0 = now_s(); }
        const cl_int reset = f <= 0 ? 1 : 0;
        const cl_uint rnd = rands[f + 1];
        CL(SetKernelArg)(k, 0, sizeof(cl_mem), &img[cur]); CL(SetKernelArg)(k, 1, sizeof(cl_mem), &img[cur ^ 1]);
        CL(SetKernelArg)(k, 9, sizeof(cl_int), &reset); CL(SetKernelArg)(k, 10, sizeof(cl_uint), &rnd);
        cl_event ev = NULL;
        err = CL(EnqueueNDRangeKernel)(q, k, 2, NULL, gws, NULL, 0, NULL, &ev); if (err) fail("clEnqueueNDRangeKernel", err);
        if (f >= 0 && (f % 8) == 0) {            /* event profiling like the reference's exec_time_rk, on a sample of the frames */
            CL(Finish)(q);
            cl_ulong a = 0, b = 0;
            CL(GetEventProfilingInfo)(ev, CL_PROFILING_COMMAND_START, sizeof a, &a, NULL); CL(GetEventProfilingInfo)(ev, CL_PROFILING_COMMAND_END, sizeof b, &b, NULL);
            kernel_ms += (double)(b - a) * 1e-6;
        }
        if (ev) CL(ReleaseEvent)(ev);
        cur ^= 1;
    }
    err = CL(Finish)(q); if (err) fail("clFinish", err);
    const double wall_s = now_s() - wall0;
    float* out = (float*)malloc((size_t)W * H * 16);
    const size_t origin[3] = {0, 0, 0}, region[3] = {(size_t)W, (size_t)H, 1};
    err = CL(EnqueueReadImage)(q, img[cur ^ 1], 1, origin, region, 0, 0, out, 0, NULL, NULL); if (err) fail("clEnqueueReadImage", err);
    char p[1024]; snprintf(p, sizeof p, "%s/image.bin", dir);
    FILE* fo = fopen(p, "wb"); if (fo) { fwrite(out, 16, (size_t)W * H, fo); fclose(fo); }
    const int n_prof = (frames + 7) / 8;
    printf("{\"impl\": \"reference-opencl\", \"platform\": \"%s\", \"device\": \"%s\", \"device_version\": \"%s\", \"bound_via\": \"%s\", \"build_options\": \"%s\", "
           "\"build_s\": %.2f, \"width\": %d, \"height\": %d, \"frames\": %d, \"ms_per_frame\": %.4f, \"kernel_ms_per_frame\": %.4f, \"msamples_s\": %.2f}\n",
           pname, dname, dver, how, opts, build_s, W, H, frames, wall_s * 1e3 / frames, n_prof ? kernel_ms / n_prof : 0.0, (double)W * H * frames / wall_s / 1e6);
    return 0;
}
