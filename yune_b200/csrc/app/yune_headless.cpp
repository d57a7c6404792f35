// yune_headless -- command-line front end of the headless renderer (replaces RendererGUI's menus, src/RendererGUI.cpp:
// "Load OBJ", kernel file, window size, GI check, "Save At Samples").  Usage:
//   yune_headless --obj scene.obj [--kernel udpt.cl|bdpt.cl] [--opts -DMIS] [--width 1024 --height 1024] [--spp 64]
//                 [--seed 12345] [--no-gi] [--bins 20] [--fov 60] [--out image.hdr|.png|.jpg|.pfm|.ppm] [--device 0]
//                 [--save-at N --save-at-out file [--save-at-ext .jpg|.png|.hdr]]     ("Save At Samples": image after N spp)
//                 [--frame-by-frame [--pipeline]]   one yune_render per sample like the reference's viewer (src/RendererCore.cpp:483-486);
//                                 --pipeline sets option "pipeline": a frame returns when its samples are handed out (yune_cuda.h)
//                 [--device-bvh [--leaf-max N]]     build the BVH on the GPU (yune_build_bvh_on_device) instead of uploading the host-built one
//                 [--option key=value ...]          any yune_set_option tunable (include/yune_cuda.h)
//                 [--gpus N]     N > 1 (0 = every device): the sample range is sharded over N devices (yune_group_*), one ncclReduce
#include "RendererCore.h"
#include "ImageIO.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

int main(int argc, char** argv)
{
    std::string obj, kernel = "udpt.cl", opts, out, save_at_out, save_at_ext;
    int save_at = 0;
    int width = 1024, height = 1024, spp = 64, bins = 20, device = 0, gpus = 1;
    unsigned seed = 12345; bool gi = true, frame_by_frame = false, pipeline = false, device_bvh = false; float fov = 60.0f; int leaf_max = 2;
    std::vector<std::pair<std::string, double>> options;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { std::cerr << "missing value for " << a << "\n"; std::exit(2); } return argv[++i]; };
        if (a == "--obj") obj = next(); else if (a == "--kernel") kernel = next(); else if (a == "--opts") opts = next();
        else if (a == "--width") width = std::atoi(next()); else if (a == "--height") height = std::atoi(next());
        else if (a == "--spp") spp = std::atoi(next()); else if (a == "--seed") seed = (unsigned)std::strtoul(next(), nullptr, 10);
        else if (a == "--bins") bins = std::atoi(next()); else if (a == "--device") device = std::atoi(next());
        else if (a == "--fov") fov = (float)std::atof(next()); else if (a == "--out") out = next(); else if (a == "--no-gi") gi = false;
        else if (a == "--gpus") gpus = std::atoi(next());
        else if (a == "--frame-by-frame") frame_by_frame = true; else if (a == "--pipeline") pipeline = true;
        else if (a == "--device-bvh") device_bvh = true; else if (a == "--leaf-max") leaf_max = std::atoi(next());
        else if (a == "--option") { std::string kv = next(); const size_t eq = kv.find('='); if (eq == std::string::npos) { std::cerr << "--option needs key=value\n"; return 2; } options.push_back({kv.substr(0, eq), std::atof(kv.c_str() + eq + 1)}); }
        else if (a == "--save-at") save_at = std::atoi(next()); else if (a == "--save-at-out") save_at_out = next(); else if (a == "--save-at-ext") save_at_ext = next();
        else { std::cerr << "unknown argument " << a << "\n"; return 2; }
    }
    if (obj.empty()) { std::cerr << "usage: yune_headless --obj scene.obj [--kernel udpt.cl] [--opts -DMIS] [--width W --height H] [--spp N] [--out image.hdr|.png|.jpg]\n"; return 2; }
    if (gpus != 1) {
        // Multi-GPU: every device renders its shard of the sample range (by sample index), one NCCL sum-reduce to device 0, which
        // tonemaps and writes the image -- the same picture as --gpus 1 up to fp32 summation order.
        try {
            yune::Scene scene;
            scene.bvh.bins = bins;
            scene.loadModel(obj, obj.substr(obj.find_last_of("/") + 1));
            if (bins > 0 && bins != 20) scene.loadBVH(bins);
            scene.main_camera.y_FOV = fov; scene.main_camera.updateViewPlaneDist();
            std::cout << "Total Triangles Loaded: " << scene.vert_data.size() << "\nBVH Size: " << scene.bvh.gpu_node_list.size() << " Nodes\n";
            yune_group* g = nullptr;
            if (yune_group_create(gpus, nullptr, &g) != YUNE_OK) { std::cerr << yune_group_last_error(nullptr) << "\n"; return 1; }
            yune::Cam cam; scene.main_camera.setBuffer(&cam);
            const bool ok = yune_group_create_render_program(g, kernel.c_str(), opts.c_str()) == YUNE_OK && yune_group_create_postproc_program(g, "tonemap.cl", "") == YUNE_OK
                && yune_group_setup_vertex_buffer(g, scene.vert_data.data(), (int)scene.vert_data.size()) == YUNE_OK
                && yune_group_setup_mat_buffer(g, scene.mat_data.data(), (int)scene.mat_data.size()) == YUNE_OK
                && yune_group_setup_bvh_buffer(g, scene.bvh.gpu_node_list.data(), (int)scene.bvh.gpu_node_list.size()) == YUNE_OK
                && yune_group_setup_image_buffers(g, width, height) == YUNE_OK && yune_group_setup_camera_buffer(g, &cam) == YUNE_OK
                && yune_group_render(g, 0, spp, gi ? 1 : 0, seed, 1) == YUNE_OK && yune_group_reduce(g, 0) == YUNE_OK;
            if (!ok) { std::cerr << yune_group_last_error(g) << "\n"; yune_group_destroy(g); return 1; }
            yune_group_stats st; yune_group_get_stats(g, &st);
            std::printf("devices %d  samples/pixel %d  render %.3f ms (slowest rank; fastest %.3f)  reduce %.3f ms  %.1f Msamples/s  %.1f Mrays/s\n", st.n_devices, spp,
                        st.render_ms_max, st.render_ms_min, st.reduce_ms, st.samples / (st.render_ms_max + st.reduce_ms) / 1e3,
                        (st.extend_rays + st.shadow_rays) / (st.render_ms_max + st.reduce_ms) / 1e3);
            int rc = 0;
            if (!out.empty()) {
                yune_ctx* root = yune_group_ctx(g, 0);
                std::vector<float> img((size_t)width * height * 4);
                const std::string ext = yune::imageExtension(out);
                const bool ldr = yune::imageIsLdr(ext);
                std::string err;
                if (ldr ? (yune_tonemap(root) != YUNE_OK || yune_read_ldr(root, img.data()) != YUNE_OK) : (yune_read_hdr(root, img.data()) != YUNE_OK)) { std::cerr << yune_last_error(root) << "\n"; rc = 1; }
                else if (!yune::writeImage(out, ext, img.data(), width, height, err)) { std::cerr << err << "\n"; rc = 1; }
                else std::cout << "wrote " << out << "\n";
            }
            yune_group_destroy(g);
            return rc;
        } catch (const std::exception& e) { std::cerr << e.what() << "\n"; return 1; }
    }
    try {
        yune::CUDAManager manager;
        manager.setup(device);                                           // throws without a B200: there is no CPU path
        manager.rk_compiler_opts = opts;
        if (!manager.createRenderProgram(kernel) || !manager.createPostProcProgram("tonemap.cl")) { std::cerr << manager.last_message << "\n"; return 1; }
        yune::RendererCore core(manager, width, height);
        core.seed = seed;
        if (save_at > 0 && !save_at_out.empty()) { core.save_at_samples = save_at; core.save_samples_fn = save_at_out; if (!save_at_ext.empty()) core.save_samples_ext = save_at_ext; }
        core.render_scene.bvh.bins = bins;
        if (!core.loadScene(obj, obj.substr(obj.find_last_of("/") + 1))) { std::cerr << manager.last_message << "\n"; return 1; }
        if (bins > 0 && bins != 20) core.render_scene.loadBVH(bins);
        core.render_scene.main_camera.y_FOV = fov; core.render_scene.main_camera.updateViewPlaneDist();
        std::cout << "Total Triangles Loaded: " << core.render_scene.vert_data.size() << "\nBVH Size: " << core.render_scene.bvh.gpu_node_list.size() << " Nodes\n";
        for (const auto& kv : options)
            if (yune_set_option(manager.ctx, kv.first.c_str(), kv.second) != YUNE_OK) { std::cerr << yune_last_error(manager.ctx) << "\n"; return 1; }
        if (!core.setup(gi)) { std::cerr << manager.last_message << "\n"; return 1; }
        if (device_bvh) {
            if (yune_build_bvh_on_device(manager.ctx, leaf_max) != YUNE_OK) { std::cerr << yune_last_error(manager.ctx) << "\n"; return 1; }
            int n_nodes = 0, n_inner = 0, depth = 0; float ms = 0;
            yune_bvh_info(manager.ctx, &n_nodes, &n_inner, &depth, &ms);
            std::printf("BVH built on the device: %d nodes (%d inner), depth %d, %.2f ms\n", n_nodes, n_inner, depth, ms);
        }
        if (frame_by_frame) {
            if (pipeline && yune_set_option(manager.ctx, "pipeline", 1) != YUNE_OK) { std::cerr << yune_last_error(manager.ctx) << "\n"; return 1; }
            if (!core.enqueueKernels(1, gi)) { std::cerr << manager.last_message << "\n"; return 1; }      // first frame: pool allocation
            const auto t0 = std::chrono::steady_clock::now();
            for (int f = 1; f < spp; f++) if (!core.enqueueKernels(1, gi)) { std::cerr << manager.last_message << "\n"; return 1; }
            const auto t1 = std::chrono::steady_clock::now();
            if (!core.finish()) { std::cerr << manager.last_message << "\n"; return 1; }
            const auto t2 = std::chrono::steady_clock::now();
            std::printf("frame by frame%s: %d frames, %.3f ms/frame (wall clock, frames 2..%d), finish %.3f ms\n", pipeline ? " (pipelined)" : "", spp,
                        spp > 1 ? std::chrono::duration<double, std::milli>(t1 - t0).count() / (spp - 1) : 0.0, spp, std::chrono::duration<double, std::milli>(t2 - t1).count());
        } else if (!core.enqueueKernels(spp, gi)) { std::cerr << manager.last_message << "\n"; return 1; }
        std::printf("samples/pixel %d  ms/frame %.4f  render time %.3f s  %.1f Msamples/s  %.1f Mrays/s  wavefront iterations %u\n",
                    core.samples_taken, core.mspf_avg, core.time_passed, core.msamples_per_s, core.mrays_per_s, core.stats.iterations);
        if (!out.empty()) { if (!core.saveImage(out)) { std::cerr << manager.last_message << "\n"; return 1; } std::cout << "wrote " << out << "\n"; }
    } catch (const std::exception& e) { std::cerr << e.what() << "\n"; return 1; }
    return 0;
}
