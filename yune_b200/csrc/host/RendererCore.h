// RendererCore.h -- headless frame scheduler: the launch protocol of yune::RendererCore (include/RendererCore.h:46-90,
// src/RendererCore.cpp:155-246 setup, :248-469 enqueueKernels, :483-505 endFrame metrics, :608-646 saveImage) without the
// GLFW/GL window, driving the C ABI of include/yune_cuda.h instead of raw cl_kernel / cl_mem handles.
#ifndef YUNE_RENDERERCORE_H
#define YUNE_RENDERERCORE_H

#include "Scene.h"
#include "yune_cuda.h"

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace yune
{
    /** Thin C++ owner of a yune_ctx with CLManager's public method set (include/CLManager.h:50-88). */
    class CUDAManager
    {
        public:
            CUDAManager();
            ~CUDAManager();
            void setup(int device = 0);                                                   /**< throws std::runtime_error like CLManager::setup */
            bool createRenderProgram(std::string fn, std::string path = "", bool reload = false);   /**< honours '#yune-preproc compiler-opts' of `path` */
            bool createPostProcProgram(std::string fn, std::string path = "", bool reload = false);
            void setupCameraBuffer(Cam* cam_data);
            bool setupImageBuffers(int width, int height);
            bool setupBVHBuffer(std::vector<BVHNodeGPU>& bvh_data, float bvh_size = 0, float scene_size = 0);
            bool setupVertexBuffer(std::vector<TriangleGPU>& vert_data, float scene_size = 0);
            bool setupMatBuffer(std::vector<Material>& mat_data);
            static void checkError(int err_code, yune_ctx* ctx, std::string filename, int line_number);   /**< negative code -> std::runtime_error */

            /** CLManager::setGuiMessageCb (include/CLManager.h:76): called as cb(message, title, log) whenever a bool method fails
             *  (src/CLManager.cpp:261-265); last_message keeps the text either way. */
            void setGuiMessageCb(std::function<void(const std::string&, const std::string&, const std::string&)> cb) { message_cb = cb; }
            void message(const std::string& msg, const std::string& title, const std::string& log = "") { last_message = msg; if (message_cb) message_cb(msg, title, log); }

            std::string rk_file, rk_compiler_opts, ppk_file, last_message;
            yune_ctx* ctx;
        private:
            bool report(int rc);
            std::function<void(const std::string&, const std::string&, const std::string&)> message_cb;
    };

    class RendererCore
    {
        public:
            RendererCore(CUDAManager& cuda_manager, int width, int height);
            bool loadScene(std::string path, std::string fn);                             /**< src/RendererCore.cpp:124-137 */
            bool setup(bool gi_check);                                                    /**< upload scene + camera, allocate images */
            bool enqueueKernels(int frames, bool new_gi_check);                           /**< `frames` more samples per pixel; blocks until done */
            bool finish();                       // option "pipeline": complete the paths still in flight (yune_finish)
        bool postProcess();
            bool reloadMatFile();                                                         /**< src/RendererCore.cpp:109-122 + the buffer update the GUI triggers (src/RendererGUI.cpp:203-220) */
            void resetValues();                                                           /**< src/RendererCore.cpp:77-86: counters and metrics back to zero; the next frame starts a new image */
            void stop();                                                                  /**< src/RendererCore.cpp:88-107: wait for the device, then resetValues() */
            void setGuiMessageCb(std::function<void(const std::string&, const std::string&, const std::string&)> cb) { cl_manager.setGuiMessageCb(cb); }
            /** src/RendererCore.cpp:608-646 without stb: ".hdr" (Radiance RGBE), ".pfm" (float RGB), ".ppm" (8-bit, tonemapped). Rows are
             *  flipped to top-down on the way out like the reference does. */
            bool saveImage(const std::string& path);                                      /**< format by extension: .hdr .png .jpg (.pfm .ppm) */
            bool saveImage(const std::string& save_fn, const std::string& save_ext);      /**< the reference's signature (src/RendererCore.cpp:608) */

            Scene render_scene;
            std::uint32_t seed;
            int samples_taken;
            float mspf_avg, ms_per_rk, ms_per_ppk, time_passed;                           /**< endFrame() metrics, :483-505 */
            double msamples_per_s, mrays_per_s;
            yune_stats stats;
            int save_at_samples;                                                          /**< "Save At Samples" (src/RendererCore.cpp:447-449): 0 = off */
            std::string save_samples_fn, save_samples_ext;                                /**< file as given; format ".jpg" (default, :57) ".png" ".hdr" */
        private:
            CUDAManager& cl_manager;
            int width, height;
            bool gi_check;
    };
}
#endif
