/* Storage-only glm::vec4 stand-in (include/Camera.h members). Oracle only. */
#ifndef YUNE_ORACLE_SHIM_GLM_VEC4
#define YUNE_ORACLE_SHIM_GLM_VEC4
namespace glm { struct vec4 { float x, y, z, w; vec4() : x(0), y(0), z(0), w(0) {} vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} }; }
#endif
