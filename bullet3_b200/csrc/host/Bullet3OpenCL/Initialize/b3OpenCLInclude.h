// Drop-in replacement for src/Bullet3OpenCL/Initialize/b3OpenCLInclude.h.
// The reference's class signatures carry OpenCL handle types (b3GpuRigidBodyPipeline.h:36,
// b3GpuNarrowPhase.h:24, b3GpuBroadphaseInterface.h:15); on the B200 build they are plain
// opaque handles: no OpenCL header, ICD or clew is involved.
//   cl_context        unused (may be 0)
//   cl_device_id      CUDA device ordinal + 1, cast to a pointer (0 = device 0)
//   cl_command_queue  a cudaStream_t (0 = the library creates a private stream)
//   cl_mem            a CUDA device pointer with the reference's AoS layout
#ifndef B3_OPENCL_INCLUDE_H
#define B3_OPENCL_INCLUDE_H
#include <stdint.h>
typedef struct b3b200_cl_context_* cl_context;
typedef struct b3b200_cl_device_* cl_device_id;
typedef void* cl_command_queue;
typedef void* cl_mem;
typedef int cl_int;
#define CL_SUCCESS 0
inline cl_device_id b3b200MakeDeviceId(int cudaOrdinal) { return (cl_device_id)(intptr_t)(cudaOrdinal + 1); }
inline int b3b200DeviceOrdinal(cl_device_id d) { return d ? (int)(intptr_t)d - 1 : 0; }
#endif
