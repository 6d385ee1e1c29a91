// Minimal b3OpenCLArray<T> (src/Bullet3OpenCL/ParallelPrimitives/b3OpenCLArray.h:7-298): a typed view of a
// device buffer owned by the B200 library, enough for the accessors of b3GpuBroadphaseInterface
// (getAllAabbsGPU, getOverlappingPairsGPU, ...): size(), getBufferCL(), copyToHost(), at().
#ifndef B3_OPENCL_ARRAY_H
#define B3_OPENCL_ARRAY_H
#include "Bullet3Common/b3AlignedObjectArray.h"
#include "Bullet3OpenCL/Initialize/b3OpenCLInclude.h"
extern "C" int b3b200_device_to_host(void* dstHost, const void* srcDevice, unsigned long long bytes, int device);
template <typename T>
class b3OpenCLArray
{
	void* m_ptr;
	int m_size;
	int m_device;

public:
	b3OpenCLArray() : m_ptr(0), m_size(0), m_device(0) {}
	void setView(void* devicePtr, int size, int device)
	{
		m_ptr = devicePtr;
		m_size = size;
		m_device = device;
	}
	cl_mem getBufferCL() const { return m_ptr; }
	int size() const { return m_size; }
	void copyToHost(b3AlignedObjectArray<T>& dst, bool = true) const
	{
		dst.resize(m_size);
		if (m_size) b3b200_device_to_host(&dst[0], m_ptr, sizeof(T) * (unsigned long long)m_size, m_device);
	}
	void copyToHostPointer(T* dst, int n, int firstElem = 0, bool = true) const
	{
		if (n > 0) b3b200_device_to_host(dst, (const char*)m_ptr + sizeof(T) * (size_t)firstElem, sizeof(T) * (unsigned long long)n, m_device);
	}
	T at(int i) const
	{
		T v;
		copyToHostPointer(&v, 1, i);
		return v;
	}
};
#endif
