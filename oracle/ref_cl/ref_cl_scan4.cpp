// ref_cl_scan4.cpp -- b3PrefixScanFloat4CL::executeHost (b3PrefixScanFloat4CL.cpp:95-120).  Its own translation unit: the
// reference header shares its include guard with b3PrefixScanCL.h (B3_PREFIX_SCAN_CL_H).  TEST INFRASTRUCTURE ONLY.
#include <string.h>
#include "Bullet3OpenCL/Initialize/b3OpenCLInclude.h"
#include "Bullet3OpenCL/ParallelPrimitives/b3PrefixScanFloat4CL.h"
#include "../../include/b3b200_types.h"

extern "C" void b3ref_cl_init();

extern "C" void refcl_prefix_scan_float4(const b3b200_float4* src, b3b200_float4* dst, int n, b3b200_float4* sum)
{
	static bool done = false;
	if (!done) b3ref_cl_init();
	done = true;
	b3PrefixScanFloat4CL scan(0, 0, 0, n + 16);
	b3AlignedObjectArray<b3Vector3> a, b;
	a.resize(n);
	b.resize(n);
	if (n) memcpy(&a[0], src, 16 * (size_t)n);
	b3Vector3 s;
	scan.executeHost(a, b, n, &s);
	if (n) memcpy(dst, &b[0], 16 * (size_t)n);
	if (sum) memcpy(sum, &s, 16);
}
