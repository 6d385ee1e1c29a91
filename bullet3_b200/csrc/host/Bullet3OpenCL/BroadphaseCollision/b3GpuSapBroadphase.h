// b3GpuSapBroadphase with the reference's surface (src/Bullet3OpenCL/BroadphaseCollision/b3GpuSapBroadphase.h:14-141)
// on top of the C ABI (include/b3b200.h: b3b200_bp_*).  b3GpuGridBroadphase shares the implementation.
#ifndef B3_GPU_SAP_BROADPHASE_H
#define B3_GPU_SAP_BROADPHASE_H
#include "b3GpuBroadphaseInterface.h"

struct b3b200_broadphase;
struct b3b200_world;

class b3B200BroadphaseBase : public b3GpuBroadphaseInterface
{
protected:
	int m_kind;
	int m_device;
	void* m_stream;
	b3b200_broadphase* m_bp;  // stand-alone handle (PairBench-style use), created lazily
	b3b200_world* m_world;    // set when a b3GpuRigidBodyPipeline adopts this broadphase
	struct Proxy
	{
		float mn[3], mx[3];
		int userPtr;
		bool large;
	};
	b3AlignedObjectArray<Proxy> m_pending;
	b3AlignedObjectArray<b3SapAabb> m_allAabbsCPU;
	b3OpenCLArray<b3SapAabb> m_allAabbsGPU;
	b3OpenCLArray<b3Int4> m_pairsGPU;
	b3OpenCLArray<int> m_smallGPU, m_largeGPU;
	int m_numOverlap;
	void ensureStandalone(int maxPairs);
	friend class b3GpuRigidBodyPipeline;

public:
	b3B200BroadphaseBase(int kind, cl_device_id device, cl_command_queue q);
	virtual ~b3B200BroadphaseBase();
	virtual void createProxy(const b3Vector3& aabbMin, const b3Vector3& aabbMax, int userPtr, int collisionFilterGroup, int collisionFilterMask);
	virtual void createLargeProxy(const b3Vector3& aabbMin, const b3Vector3& aabbMax, int userPtr, int collisionFilterGroup, int collisionFilterMask);
	virtual void calculateOverlappingPairs(int maxPairs);
	// the reference's host twin is an O(N^2) CPU loop; there is no CPU path here: same result from the device
	virtual void calculateOverlappingPairsHost(int maxPairs) { calculateOverlappingPairs(maxPairs); }
	virtual void writeAabbsToGpu();
	virtual cl_mem getAabbBufferWS();
	virtual int getNumOverlap();
	virtual cl_mem getOverlappingPairBuffer();
	virtual b3OpenCLArray<b3SapAabb>& getAllAabbsGPU();
	virtual b3AlignedObjectArray<b3SapAabb>& getAllAabbsCPU();
	virtual b3OpenCLArray<b3Int4>& getOverlappingPairsGPU();
	virtual b3OpenCLArray<int>& getSmallAabbIndicesGPU() { return m_smallGPU; }
	virtual b3OpenCLArray<int>& getLargeAabbIndicesGPU() { return m_largeGPU; }
	void reset();
};

class b3GpuSapBroadphase : public b3B200BroadphaseBase
{
public:
	enum b3GpuSapKernelType
	{
		B3_GPU_SAP_KERNEL_BRUTE_FORCE_CPU = 1,
		B3_GPU_SAP_KERNEL_BRUTE_FORCE_GPU,
		B3_GPU_SAP_KERNEL_ORIGINAL,
		B3_GPU_SAP_KERNEL_BARRIER,
		B3_GPU_SAP_KERNEL_LOCAL_SHARED_MEMORY
	};
	// all kernel variants of the reference compute the same pair set; one sm_100a sweep serves them all
	b3GpuSapBroadphase(cl_context, cl_device_id device, cl_command_queue q, b3GpuSapKernelType = B3_GPU_SAP_KERNEL_LOCAL_SHARED_MEMORY)
		: b3B200BroadphaseBase(0 /*B3B200_BP_SAP*/, device, q) {}
	static b3GpuBroadphaseInterface* CreateFuncBruteForceCpu(cl_context ctx, cl_device_id device, cl_command_queue q) { return new b3GpuSapBroadphase(ctx, device, q, B3_GPU_SAP_KERNEL_BRUTE_FORCE_CPU); }
	static b3GpuBroadphaseInterface* CreateFuncBruteForceGpu(cl_context ctx, cl_device_id device, cl_command_queue q) { return new b3GpuSapBroadphase(ctx, device, q, B3_GPU_SAP_KERNEL_BRUTE_FORCE_GPU); }
	static b3GpuBroadphaseInterface* CreateFuncOriginal(cl_context ctx, cl_device_id device, cl_command_queue q) { return new b3GpuSapBroadphase(ctx, device, q, B3_GPU_SAP_KERNEL_ORIGINAL); }
	static b3GpuBroadphaseInterface* CreateFuncBarrier(cl_context ctx, cl_device_id device, cl_command_queue q) { return new b3GpuSapBroadphase(ctx, device, q, B3_GPU_SAP_KERNEL_BARRIER); }
	static b3GpuBroadphaseInterface* CreateFuncLocalMemory(cl_context ctx, cl_device_id device, cl_command_queue q) { return new b3GpuSapBroadphase(ctx, device, q, B3_GPU_SAP_KERNEL_LOCAL_SHARED_MEMORY); }
};
#endif
