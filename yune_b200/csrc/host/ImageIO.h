// ImageIO.h -- image export (the stb_image_write calls of RendererCore::saveImage, src/RendererCore.cpp:608-646).
#ifndef YUNE_IMAGE_IO_H
#define YUNE_IMAGE_IO_H

#include <string>

namespace yune {

/** Lower-case extension of `path` including the dot ("" when there is none). */
std::string imageExtension(const std::string& path);
/** ".png" ".jpg" ".jpeg" ".ppm": written from the tonemapped image, 8 bits per channel. */
bool imageIsLdr(const std::string& ext);
/** ".hdr" ".pfm": written from the float image. */
bool imageIsHdr(const std::string& ext);
/** Encode a bottom-up RGBA float image (the layout of yune_read_hdr / yune_read_ldr) into `path`.  The file's first row is
 *  the top of the picture (stbi_flip_vertically_on_write(1) in the reference); alpha is dropped. */
bool writeImage(const std::string& path, const std::string& ext, const float* rgba, int width, int height, std::string& err);

}
#endif
