/* Minimal stand-in for <CL/cl.h>, just enough for the reference's Whitted host
 * sources (common.h, scene.c, bitmap.c) to compile without an OpenCL SDK.
 * TEST INFRASTRUCTURE ONLY (used by oracle/Makefile to build oracle/_ref).
 *
 * cl_float4 / cl_uchar4 are deliberately NOT 16-byte aligned: the reference was
 * built against a 2011 cl_platform.h whose CL_ALIGNED was empty, which is what
 * makes its nested `Primitive` 96 bytes (last row). */
#ifndef ORACLE_SHIM_CL_H
#define ORACLE_SHIM_CL_H
#include <stdint.h>
typedef float    cl_float;
typedef uint32_t cl_uint;
typedef int32_t  cl_int;
typedef uint32_t cl_bool;
typedef uint8_t  cl_uchar;
typedef union { cl_float s[4]; } cl_float4;
typedef union { cl_uchar s[4]; } cl_uchar4;
typedef struct _shim_cl_mem *cl_mem;
typedef struct _shim_cl_context *cl_context;
typedef struct _shim_cl_device_id *cl_device_id;
typedef struct _shim_cl_command_queue *cl_command_queue;
typedef struct _shim_cl_program *cl_program;
typedef struct _shim_cl_kernel *cl_kernel;
#endif
