/* Stub <CL/cl.h> -- test infrastructure only.
 *
 * The reference's ViT_seq.c (line 8) and kernelHandler.h (line 12, pulled in by
 * ViT_opencl.h) include <CL/cl.h> although the CPU path never calls OpenCL.
 * This stub supplies just the names those headers mention so the reference
 * sources compile UNMODIFIED with gcc on a box without OpenCL headers.
 */
#ifndef VITB200_STUB_CL_H
#define VITB200_STUB_CL_H
#include <stddef.h>
#include <stdio.h>
typedef int cl_int;
typedef unsigned int cl_uint;
typedef struct _stub_cl_program *cl_program;
typedef struct _stub_cl_device_id *cl_device_id;
#define CL_SUCCESS 0
#endif
