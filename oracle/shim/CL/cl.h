/* Minimal stand-in for <CL/cl.h>, just enough for the reference's HOST sources
 * (include/CL_headers.h, src/Scene.cpp, src/BVH.cpp, ...) to compile in place.
 * TEST INFRASTRUCTURE ONLY (oracle/): never included by the product path.
 * The reference only uses the scalar typedefs and cl_float4 with its .s[4] view. */
#ifndef YUNE_ORACLE_SHIM_CL_H
#define YUNE_ORACLE_SHIM_CL_H
#include <stdint.h>
typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef float    cl_float;
typedef union alignas(16) cl_float4_u { cl_float s[4]; } cl_float4;
typedef unsigned int GLuint;
#endif
