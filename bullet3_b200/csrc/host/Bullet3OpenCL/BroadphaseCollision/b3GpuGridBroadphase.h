// b3GpuGridBroadphase (src/Bullet3OpenCL/BroadphaseCollision/b3GpuGridBroadphase.h:7-78)
#ifndef B3_GPU_GRID_BROADPHASE_H
#define B3_GPU_GRID_BROADPHASE_H
#include "b3GpuSapBroadphase.h"
class b3GpuGridBroadphase : public b3B200BroadphaseBase
{
public:
	b3GpuGridBroadphase(cl_context, cl_device_id device, cl_command_queue q) : b3B200BroadphaseBase(1 /*B3B200_BP_GRID*/, device, q) {}
	static b3GpuBroadphaseInterface* CreateFunc(cl_context ctx, cl_device_id device, cl_command_queue q) { return new b3GpuGridBroadphase(ctx, device, q); }
};
#endif
