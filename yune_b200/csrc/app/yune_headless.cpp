// yune_headless -- command-line front end of the headless renderer (replaces RendererGUI's menus, src/RendererGUI.cpp:
// "Load OBJ", kernel file, window size, GI check, "Save At Samples").  Usage:
//   yune_headless --obj scene.obj [--kernel udpt.cl|bdpt.cl] [--opts -DMIS] [--width 1024 --height 1024] [--spp 64]
//                 [--seed 12345] [--no-gi] [--bins 20] [--fov 60] [--out image.hdr|.png|.jpg|.pfm|.ppm] [--device 0]
//                 [--save-at N --save-at-out file [--save-at-ext .jpg|.png|.hdr]]     ("Save At Samples": image after N spp)
#include "RendererCore.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

int main(int argc, char** argv)
{
    std::string obj, kernel = "udpt.cl", opts, out, save_at_out, save_at_ext;
    int save_at = 0;
    int width = 1024, height = 1024, spp = 64, bins = 20, device = 0;
    unsigned seed = 12345; bool gi = true; float fov = 60.0f;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { std::cerr << "missing value for " << a << "\n"; std::exit(2); } return argv[++i]; };
        if (a == "--obj") obj = next(); else if (a == "--kernel") kernel = next(); else if (a == "--opts") opts = next();
        else if (a == "--width") width = std::atoi(next()); else if (a == "--height") height = std::atoi(next());
        else if (a == "--spp") spp = std::atoi(next()); else if (a == "--seed") seed = (unsigned)std::strtoul(next(), nullptr, 10);
        else if (a == "--bins") bins = std::atoi(next()); else if (a == "--device") device = std::atoi(next());
        else if (a == "--fov") fov = (float)std::atof(next()); else if (a == "--out") out = next(); else if (a == "--no-gi") gi = false;
        else if (a == "--save-at") save_at = std::atoi(next()); else if (a == "--save-at-out") save_at_out = next(); else if (a == "--save-at-ext") save_at_ext = next();
        else { std::cerr << "unknown argument " << a << "\n"; return 2; }
    }
    if (obj.empty()) { std::cerr << "usage: yune_headless --obj scene.obj [--kernel udpt.cl] [--opts -DMIS] [--width W --height H] [--spp N] [--out image.hdr|.png|.jpg]\n"; return 2; }
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
        if (!core.setup(gi) || !core.enqueueKernels(spp, gi)) { std::cerr << manager.last_message << "\n"; return 1; }
        std::printf("samples/pixel %d  ms/frame %.4f  render time %.3f s  %.1f Msamples/s  %.1f Mrays/s  wavefront iterations %u\n",
                    core.samples_taken, core.mspf_avg, core.time_passed, core.msamples_per_s, core.mrays_per_s, core.stats.iterations);
        if (!out.empty()) { if (!core.saveImage(out)) { std::cerr << manager.last_message << "\n"; return 1; } std::cout << "wrote " << out << "\n"; }
    } catch (const std::exception& e) { std::cerr << e.what() << "\n"; return 1; }
    return 0;
}
