/* Storage-only glm::vec3 stand-in (src/Scene.cpp:247-248 uses .x/.y/.z only). Oracle only. */
#ifndef YUNE_ORACLE_SHIM_GLM_VEC3
#define YUNE_ORACLE_SHIM_GLM_VEC3
namespace glm { struct vec3 { float x, y, z; vec3() : x(0), y(0), z(0) {} vec3(float a, float b, float c) : x(a), y(b), z(c) {} }; }
#endif
