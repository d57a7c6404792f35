#include "RendererCore.h"
#include "ImageIO.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace yune
{
    CUDAManager::CUDAManager() : ctx(nullptr) {}
    CUDAManager::~CUDAManager() { if (ctx) yune_destroy(ctx); }

    void CUDAManager::checkError(int err_code, yune_ctx* c, std::string filename, int line_number)
    {
        if (err_code >= 0) return;
        std::ostringstream ss;
        ss << yune_last_error(c) << " in File: " << filename << " at Line number: " << line_number;   // src/CLManager.cpp:755-855
        throw std::runtime_error(ss.str());
    }
    bool CUDAManager::report(int rc)
    {
        if (rc == YUNE_OK) return true;
        message(yune_last_error(ctx), "Error!");      // the reference pushes this to the GUI callback (src/CLManager.cpp:261-265)
        return false;
    }
    void CUDAManager::setup(int device)
    {
        if (yune_setup(device, &ctx) != YUNE_OK) throw std::runtime_error(yune_last_error(nullptr));
    }
    bool CUDAManager::createRenderProgram(std::string fn, std::string path, bool)
    {
        // '#yune-preproc compiler-opts <one token>' (src/CLManager.cpp:182-204)
        if (!path.empty()) {
            std::ifstream f(path);
            std::string line, a, b, c;
            while (f.is_open() && std::getline(f, line)) {
                std::istringstream ss(line);
                if ((ss >> a >> b >> c) && a == "#yune-preproc" && b == "compiler-opts") rk_compiler_opts = c;
            }
        }
        if (!report(yune_create_render_program(ctx, fn.c_str(), rk_compiler_opts.c_str()))) return false;
        rk_file = fn;
        return true;
    }
    bool CUDAManager::createPostProcProgram(std::string fn, std::string, bool)
    {
        if (!report(yune_create_postproc_program(ctx, fn.c_str(), ""))) return false;
        ppk_file = fn;
        return true;
    }
    void CUDAManager::setupCameraBuffer(Cam* cam) { checkError(yune_setup_camera_buffer(ctx, cam), ctx, __FILE__, __LINE__); }
    bool CUDAManager::setupImageBuffers(int w, int h) { return report(yune_setup_image_buffers(ctx, w, h)); }
    bool CUDAManager::setupBVHBuffer(std::vector<BVHNodeGPU>& v, float, float) { return report(yune_setup_bvh_buffer(ctx, v.data(), (int)v.size())); }
    bool CUDAManager::setupVertexBuffer(std::vector<TriangleGPU>& v, float) { return report(yune_setup_vertex_buffer(ctx, v.data(), (int)v.size())); }
    bool CUDAManager::setupMatBuffer(std::vector<Material>& v) { return report(yune_setup_mat_buffer(ctx, v.data(), (int)v.size())); }

    RendererCore::RendererCore(CUDAManager& m, int w, int h)
        : seed(12345), samples_taken(0), mspf_avg(0), ms_per_rk(0), ms_per_ppk(0), time_passed(0), msamples_per_s(0), mrays_per_s(0),
          save_at_samples(0), save_samples_ext(".jpg"), cl_manager(m), width(w), height(h), gi_check(true)
    {
        std::memset(&stats, 0, sizeof(stats));
    }

    bool RendererCore::loadScene(std::string path, std::string fn)
    {
        try { render_scene.loadModel(path, fn); }
        catch (const std::exception& e) { cl_manager.message(e.what(), "Error loading File!"); return false; }
        return true;
    }

    bool RendererCore::reloadMatFile()
    {
        try { render_scene.reloadMatFile(); }
        catch (const std::exception& e) { cl_manager.message(e.what(), "Error loading File!"); return false; }
        if (!cl_manager.setupMatBuffer(render_scene.mat_data)) return false;
        samples_taken = 0;                                     // new materials: the accumulated image is stale
        return true;
    }

    void RendererCore::resetValues()
    {
        gi_check = true;
        samples_taken = 0; save_at_samples = 0; time_passed = 0;
        mspf_avg = ms_per_ppk = ms_per_rk = 0;
        msamples_per_s = mrays_per_s = 0;
    }

    void RendererCore::stop()
    {
        if (cl_manager.ctx) yune_synchronize(cl_manager.ctx);
        resetValues();
    }

    bool RendererCore::setup(bool gi)
    {
        if (!cl_manager.setupVertexBuffer(render_scene.vert_data, render_scene.scene_size_mb)) return false;
        if (!cl_manager.setupMatBuffer(render_scene.mat_data)) return false;
        if (!cl_manager.setupBVHBuffer(render_scene.bvh.gpu_node_list, render_scene.bvh.bvh_size_mb, render_scene.scene_size_mb)) return false;
        if (!cl_manager.setupImageBuffers(width, height)) return false;
        Cam cam;
        render_scene.main_camera.setBuffer(&cam);
        cl_manager.setupCameraBuffer(&cam);
        gi_check = gi; samples_taken = 0; time_passed = 0;
        return true;
    }

    bool RendererCore::enqueueKernels(int frames, bool new_gi_check)
    {
        bool reset = samples_taken == 0;
        if (render_scene.main_camera.is_changed) {                      // src/RendererCore.cpp:531-553
            Cam cam; render_scene.main_camera.setBuffer(&cam); cl_manager.setupCameraBuffer(&cam); reset = true;
        }
        if (new_gi_check != gi_check) { gi_check = new_gi_check; reset = true; }      // :556-565
        if (reset) samples_taken = 0;
        // "Save At Samples" (src/RendererCore.cpp:447-449): when the count passes save_at_samples inside this batch, the batch is
        // rendered in two calls with the save in between -- the image written holds exactly that many samples per pixel.
        bool ok = true;
        double ms = 0, samples = 0, rays = 0;
        int left = frames;
        do {
            int n = left;
            if (save_at_samples > 0 && samples_taken < save_at_samples && samples_taken + left > save_at_samples) n = save_at_samples - samples_taken;
            if (yune_render(cl_manager.ctx, samples_taken, n, gi_check ? 1 : 0, seed, reset ? 1 : 0) != YUNE_OK) {
                cl_manager.last_message = yune_last_error(cl_manager.ctx);
                return false;
            }
            reset = false;
            const bool reached = save_at_samples > 0 && n > 0 && samples_taken < save_at_samples && samples_taken + n == save_at_samples;
            samples_taken += n; left -= n;
            yune_get_stats(cl_manager.ctx, &stats);
            ms += stats.render_ms; samples += (double)stats.samples; rays += (double)(stats.extend_rays + stats.shadow_rays);
            if (reached) ok = saveImage(save_samples_fn, save_samples_ext) && ok;      // (finishes paths in flight first)
        } while (left > 0);
        // endFrame() metrics (:483-505): one frame = one sample per pixel
        mspf_avg = frames > 0 ? (float)(ms / frames) : 0.0f;
        ms_per_rk = mspf_avg;
        time_passed += (float)(ms / 1000.0);
        msamples_per_s = ms > 0 ? samples / ms / 1e3 : 0;
        mrays_per_s = ms > 0 ? rays / ms / 1e3 : 0;
        return ok;
    }

    // Option "pipeline" (yune_cuda.h): complete the paths the last frames left in flight.  A no-op otherwise.
    bool RendererCore::finish()
    {
        if (yune_finish(cl_manager.ctx) != YUNE_OK) { cl_manager.last_message = yune_last_error(cl_manager.ctx); return false; }
        yune_get_stats(cl_manager.ctx, &stats);
        return true;
    }

    bool RendererCore::postProcess()
    {
        if (yune_tonemap(cl_manager.ctx) != YUNE_OK) { cl_manager.last_message = yune_last_error(cl_manager.ctx); return false; }
        yune_get_stats(cl_manager.ctx, &stats);
        ms_per_ppk = (float)stats.tonemap_ms;
        return true;
    }

    bool RendererCore::saveImage(const std::string& path) { return saveImage(path, imageExtension(path)); }

    // RendererCore::saveImage(save_fn, save_ext) (src/RendererCore.cpp:608-646): ".hdr" from the float image, ".png" / ".jpg" from
    // the 8-bit view of the tonemapped one (the reference reads COLOR_ATTACHMENT3, the post-processing output).
    bool RendererCore::saveImage(const std::string& save_fn, const std::string& save_ext)
    {
        std::vector<float> img((size_t)width * height * 4);
        const bool ldr = imageIsLdr(save_ext);
        if (!ldr && !imageIsHdr(save_ext)) { cl_manager.last_message = "unsupported image extension (use .hdr, .png, .jpg, .pfm or .ppm)"; return false; }
        if (!finish()) return false;                                     // a saved image holds complete samples only
        if (ldr && !postProcess()) return false;
        const int rc = ldr ? yune_read_ldr(cl_manager.ctx, img.data()) : yune_read_hdr(cl_manager.ctx, img.data());
        if (rc != YUNE_OK) { cl_manager.last_message = yune_last_error(cl_manager.ctx); return false; }
        std::string err;
        if (!writeImage(save_fn, save_ext, img.data(), width, height, err)) { cl_manager.last_message = err; return false; }
        return true;
    }
}
