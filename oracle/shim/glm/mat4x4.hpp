/* Storage-only glm::mat4x4 stand-in (include/Camera.h member). Oracle only. */
#ifndef YUNE_ORACLE_SHIM_GLM_MAT4
#define YUNE_ORACLE_SHIM_GLM_MAT4
#include "vec4.hpp"
namespace glm { struct mat4x4 { vec4 c[4]; }; typedef mat4x4 mat4; }
#endif
