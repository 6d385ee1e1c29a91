// fake_cl.cpp -- a host-memory stand-in for the OpenCL runtime, so that the UNMODIFIED
// reference sources of src/Bullet3OpenCL (compiled with -DB3_USE_CLEW) can run their own HOST
// TWINS (the code paths behind gCpu* / CHECK_ON_HOST / executeHost) in this container, which has
// no OpenCL ICD.  TEST INFRASTRUCTURE ONLY.
//   * buffers are malloc'ed host memory; read/write/copy are memcpy
//   * programs/kernels are dummy handles; launching one aborts (a host twin must never do that)
// The clew function-pointer table (src/clew/clew.c, compiled unmodified) is filled by b3ref_cl_init().
// b3OpenCLUtils (context creation, JIT + binary cache: pure runtime plumbing) is replaced by the
// few C functions below instead of being compiled.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "clew/clew.h"
#include "Bullet3OpenCL/Initialize/b3OpenCLUtils.h"

int g_fakeClLaunchIsNoop = 0;
int g_fakeClLaunches = 0;

namespace
{
struct FakeMem
{
	size_t size;
	char* data;
};
int g_dummy[8];

cl_mem CL_API_CALL fCreateBuffer(cl_context, cl_mem_flags, size_t size, void* host, cl_int* err)
{
	FakeMem* m = (FakeMem*)malloc(sizeof(FakeMem));
	m->size = size;
	m->data = (char*)calloc(size ? size : 1, 1);
	if (host) memcpy(m->data, host, size);
	if (err) *err = CL_SUCCESS;
	return (cl_mem)m;
}
cl_int CL_API_CALL fReleaseMemObject(cl_mem mem)
{
	FakeMem* m = (FakeMem*)mem;
	if (m)
	{
		free(m->data);
		free(m);
	}
	return CL_SUCCESS;
}
cl_int CL_API_CALL fEnqueueReadBuffer(cl_command_queue, cl_mem mem, cl_bool, size_t off, size_t cb, void* dst, cl_uint, const cl_event*, cl_event*)
{
	memcpy(dst, ((FakeMem*)mem)->data + off, cb);
	return CL_SUCCESS;
}
cl_int CL_API_CALL fEnqueueWriteBuffer(cl_command_queue, cl_mem mem, cl_bool, size_t off, size_t cb, const void* src, cl_uint, const cl_event*, cl_event*)
{
	memcpy(((FakeMem*)mem)->data + off, src, cb);
	return CL_SUCCESS;
}
cl_int CL_API_CALL fEnqueueCopyBuffer(cl_command_queue, cl_mem src, cl_mem dst, size_t so, size_t dof, size_t cb, cl_uint, const cl_event*, cl_event*)
{
	memmove(((FakeMem*)dst)->data + dof, ((FakeMem*)src)->data + so, cb);
	return CL_SUCCESS;
}
cl_int CL_API_CALL fFinish(cl_command_queue) { return CL_SUCCESS; }
cl_int CL_API_CALL fFlush(cl_command_queue) { return CL_SUCCESS; }
cl_int CL_API_CALL fReleaseKernel(cl_kernel) { return CL_SUCCESS; }
cl_int CL_API_CALL fReleaseProgram(cl_program) { return CL_SUCCESS; }
cl_int CL_API_CALL fSetKernelArg(cl_kernel, cl_uint, size_t, const void*) { return CL_SUCCESS; }
cl_int CL_API_CALL fEnqueueNDRangeKernel(cl_command_queue, cl_kernel, cl_uint, const size_t*, const size_t*, const size_t*, cl_uint, const cl_event*, cl_event*)
{
	// GpuSatCollision::computeConvexConvexContactsGPUSAT built with CHECK_ON_HOST runs its host contact loop and
	// then still enqueues the device kernels (b3ConvexHullContact.cpp:2783-4408); with zero-filled "device" flags
	// they have nothing to add, so the narrowphase wrapper turns launches into counted no-ops.  Everywhere else a
	// launch means a host twin was not taken: abort loudly.
	if (g_fakeClLaunchIsNoop)
	{
		g_fakeClLaunches++;
		return CL_SUCCESS;
	}
	fprintf(stderr, "fake_cl: a device kernel launch was attempted -- only host twins may run in the reference build\n");
	abort();
	return CL_SUCCESS;
}
cl_int CL_API_CALL fGetDeviceInfo(cl_device_id, cl_device_info, size_t sz, void* dst, size_t* ret)
{
	if (dst && sz) memset(dst, 0, sz);
	if (ret) *ret = sz;
	return CL_SUCCESS;
}
cl_int CL_API_CALL fGetMemObjectInfo(cl_mem mem, cl_mem_info, size_t sz, void* dst, size_t* ret)
{
	if (dst && sz >= sizeof(size_t)) *(size_t*)dst = ((FakeMem*)mem)->size;
	if (ret) *ret = sizeof(size_t);
	return CL_SUCCESS;
}
}  // namespace

extern "C" {

void b3ref_cl_init()
{
	__clewCreateBuffer = fCreateBuffer;
	__clewReleaseMemObject = fReleaseMemObject;
	__clewEnqueueReadBuffer = fEnqueueReadBuffer;
	__clewEnqueueWriteBuffer = fEnqueueWriteBuffer;
	__clewEnqueueCopyBuffer = fEnqueueCopyBuffer;
	__clewFinish = fFinish;
	__clewFlush = fFlush;
	__clewReleaseKernel = fReleaseKernel;
	__clewReleaseProgram = fReleaseProgram;
	__clewSetKernelArg = fSetKernelArg;
	__clewEnqueueNDRangeKernel = fEnqueueNDRangeKernel;
	__clewGetDeviceInfo = fGetDeviceInfo;
	__clewGetMemObjectInfo = fGetMemObjectInfo;
}

// ---- replacements for src/Bullet3OpenCL/Initialize/b3OpenCLUtils.cpp (runtime plumbing only)
cl_program b3OpenCLUtils_compileCLProgramFromString(cl_context, cl_device_id, const char*, cl_int* pErrNum, const char*, const char*, bool)
{
	if (pErrNum) *pErrNum = CL_SUCCESS;
	return (cl_program)&g_dummy[0];
}
cl_kernel b3OpenCLUtils_compileCLKernelFromString(cl_context, cl_device_id, const char*, const char*, cl_int* pErrNum, cl_program, const char*)
{
	if (pErrNum) *pErrNum = CL_SUCCESS;
	return (cl_kernel)&g_dummy[1];
}
void b3OpenCLUtils_getDeviceInfo(cl_device_id, b3OpenCLDeviceInfo* info)
{
	memset(info, 0, sizeof(*info));
	info->m_deviceType = CL_DEVICE_TYPE_GPU;
	info->m_workgroupSize = 256;
	info->m_computeUnits = 1;
}
void b3OpenCLUtils_printDeviceInfo(cl_device_id) {}
int b3OpenCLUtils_clewInit() { return 0; }
int b3OpenCLUtils_getNumPlatforms(cl_int*) { return 1; }
}

// the two non-inline members of struct b3OpenCLUtils (b3OpenCLUtils.h:136,165)
void b3OpenCLUtils::getDeviceInfo(cl_device_id device, b3OpenCLDeviceInfo* info) { b3OpenCLUtils_getDeviceInfo(device, info); }
void b3OpenCLUtils::getPlatformInfo(cl_platform_id, b3OpenCLPlatformInfo* info) { *info = b3OpenCLPlatformInfo(); }
