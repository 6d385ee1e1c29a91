// common.cuh -- device math with the reference's operation order, error
// plumbing and small RAII helpers shared by every translation unit.
//
// All FP32 arithmetic on the parity-critical path is written so that, with FMA
// contraction disabled (nvcc --fmad=false), it evaluates the same IEEE-754
// operations in the same order as the reference's scalar C++ path
// (src/Bullet3Common/b3Vector3.h, b3Quaternion.h:305-310,651-729,868-879,
//  b3Matrix3x3.h:201-262, b3Transform.h:90-93).  That is what makes broadphase
// pair sets and per-pair contact counts bit-exact against the CPU oracle.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <stdio.h>
#include <string>
#include <atomic>
#include "../../include/b3b200.h"

namespace b3b200
{
// ---------------------------------------------------------------- errors
void setLastError(const char* fmt, ...);
extern std::atomic<long long> g_launchCount;  // process-wide statistic (worlds may be stepped from several threads)
extern thread_local long long g_allocEpoch;   // bumped by every device (re)allocation of THIS thread: a step graph whose capture saw one is discarded

#define B3_CUDA_CHECK(expr)                                                                 \
	do                                                                                      \
	{                                                                                       \
		cudaError_t _e = (expr);                                                            \
		if (_e != cudaSuccess)                                                              \
		{                                                                                   \
			b3b200::setLastError("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
								 __FILE__, __LINE__);                                       \
			return B3B200_ERR_CUDA;                                                         \
		}                                                                                   \
	} while (0)

#define B3_LAUNCH_CHECK()                                                                   \
	do                                                                                      \
	{                                                                                       \
		b3b200::g_launchCount++;                                                            \
		cudaError_t _e = cudaGetLastError();                                                \
		if (_e != cudaSuccess)                                                              \
		{                                                                                   \
			b3b200::setLastError("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
								 __FILE__, __LINE__);                                       \
			return B3B200_ERR_CUDA;                                                         \
		}                                                                                   \
	} while (0)

#define B3_TRY(expr)            \
	do                          \
	{                           \
		int _r = (expr);        \
		if (_r < 0) return _r;  \
	} while (0)

// ---------------------------------------------------------------- device buffer
template <typename T>
struct DevBuf
{
	T* ptr = nullptr;
	size_t cap = 0;  // elements
	~DevBuf() { release(); }
	DevBuf() {}
	DevBuf(const DevBuf&) = delete;
	DevBuf& operator=(const DevBuf&) = delete;
	void release()
	{
		if (ptr) cudaFree(ptr);
		ptr = nullptr;
		cap = 0;
	}
	// grow-only; contents are NOT preserved
	int reserve(size_t n)
	{
		if (n <= cap) return 0;
		release();
		if (n == 0) return 0;
		g_allocEpoch++;
		cudaError_t e = cudaMalloc((void**)&ptr, n * sizeof(T));
		if (e != cudaSuccess)
		{
			ptr = nullptr;
			setLastError("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
			return B3B200_ERR_CUDA;
		}
		cap = n;
		return 0;
	}
};

inline int divUp(long long a, int b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- math
#define B3_HD __host__ __device__ __forceinline__
#define B3_D __device__ __forceinline__

B3_HD float4 mk4(float x, float y, float z, float w = 0.f) { return make_float4(x, y, z, w); }
B3_HD float dot3(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
B3_HD float4 cross3(const float4& a, const float4& b)
{
	return mk4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.f);
}
B3_HD float4 add3(const float4& a, const float4& b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, 0.f); }
B3_HD float4 sub3(const float4& a, const float4& b) { return mk4(a.x - b.x, a.y - b.y, a.z - b.z, 0.f); }
B3_HD float4 scale3(const float4& a, float s) { return mk4(a.x * s, a.y * s, a.z * s, 0.f); }
B3_HD float4 neg3(const float4& a) { return mk4(-a.x, -a.y, -a.z, 0.f); }

// b3Vector3::normalized(): v * (1 / sqrt(dot)) (b3Vector3.h:802,904)
B3_HD float4 normalized3(const float4& v)
{
	float s = 1.0f / sqrtf(dot3(v, v));
	return mk4(v.x * s, v.y * s, v.z * s, 0.f);
}

// q * v  (b3Quaternion.h:724-729)
B3_HD float4 quatMulVec(const float4& q, const float4& w)
{
	return mk4(q.w * w.x + q.y * w.z - q.z * w.y,
			   q.w * w.y + q.z * w.x - q.x * w.z,
			   q.w * w.z + q.x * w.y - q.y * w.x,
			   -q.x * w.x - q.y * w.y - q.z * w.z);
}
// a *= b  (b3Quaternion.h:305-310)
B3_HD float4 quatMul(const float4& a, const float4& b)
{
	return mk4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
			   a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
			   a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
			   a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
B3_HD float4 quatInverse(const float4& q) { return mk4(-q.x, -q.y, -q.z, q.w); }
// b3QuatRotate (b3Quaternion.h:868-879)
B3_HD float4 quatRotate(const float4& q, const float4& v)
{
	float4 t = quatMulVec(q, v);
	float4 r = quatMul(t, quatInverse(q));
	return mk4(r.x, r.y, r.z, 0.f);
}

struct Mat3
{
	float4 r0, r1, r2;
};
// b3Matrix3x3::setRotation scalar path (b3Matrix3x3.h:201-262)
B3_HD Mat3 matFromQuat(const float4& q)
{
	float d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
	float s = 2.0f / d;
	float xs = q.x * s, ys = q.y * s, zs = q.z * s;
	float wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
	float xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
	float yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
	Mat3 m;
	m.r0 = mk4(1.0f - (yy + zz), xy - wz, xz + wy, 0.f);
	m.r1 = mk4(xy + wz, 1.0f - (xx + zz), yz - wx, 0.f);
	m.r2 = mk4(xz - wy, yz + wx, 1.0f - (xx + yy), 0.f);
	return m;
}
// b3Matrix3x3::getRotation scalar path (b3Matrix3x3.h:462-490)
B3_HD float4 quatFromMat(const Mat3& m)
{
	const float el[3][3] = {{m.r0.x, m.r0.y, m.r0.z}, {m.r1.x, m.r1.y, m.r1.z}, {m.r2.x, m.r2.y, m.r2.z}};
	float trace = el[0][0] + el[1][1] + el[2][2];
	float t0, t1, t2, t3;
	if (trace > 0.0f)
	{
		float s = sqrtf(trace + 1.0f);
		t3 = (s * 0.5f);
		s = 0.5f / s;
		t0 = ((el[2][1] - el[1][2]) * s);
		t1 = ((el[0][2] - el[2][0]) * s);
		t2 = ((el[1][0] - el[0][1]) * s);
	}
	else if (el[0][0] < el[1][1] ? !(el[1][1] < el[2][2]) : false)
	{
		// i = 1, j = 2, k = 0
		float s = sqrtf(el[1][1] - el[2][2] - el[0][0] + 1.0f);
		t1 = s * 0.5f;
		s = 0.5f / s;
		t3 = (el[0][2] - el[2][0]) * s;
		t2 = (el[2][1] + el[1][2]) * s;
		t0 = (el[0][1] + el[1][0]) * s;
	}
	else if (el[0][0] < el[1][1] ? true : (el[0][0] < el[2][2]))
	{
		// i = 2, j = 0, k = 1
		float s = sqrtf(el[2][2] - el[0][0] - el[1][1] + 1.0f);
		t2 = s * 0.5f;
		s = 0.5f / s;
		t3 = (el[1][0] - el[0][1]) * s;
		t0 = (el[0][2] + el[2][0]) * s;
		t1 = (el[1][2] + el[2][1]) * s;
	}
	else
	{
		// i = 0, j = 1, k = 2
		float s = sqrtf(el[0][0] - el[1][1] - el[2][2] + 1.0f);
		t0 = s * 0.5f;
		s = 0.5f / s;
		t3 = (el[2][1] - el[1][2]) * s;
		t1 = (el[1][0] + el[0][1]) * s;
		t2 = (el[2][0] + el[0][2]) * s;
	}
	return mk4(t0, t1, t2, t3);
}
B3_HD float4 matMulVec(const Mat3& m, const float4& v)
{
	return mk4(dot3(m.r0, v), dot3(m.r1, v), dot3(m.r2, v), 0.f);
}
// b3TransformPoint, C++ path (shared/b3Quat.h:18-24): basis(q) * p + t
B3_HD float4 transformPoint(const float4& p, const float4& t, const float4& q)
{
	Mat3 m = matFromQuat(q);
	float4 r = matMulVec(m, p);
	return mk4(r.x + t.x, r.y + t.y, r.z + t.z, 0.f);
}

// ---------------------------------------------------------------- loads
B3_D float4 ldg4(const float4* p) { return __ldg(p); }
B3_D int laneId() { return threadIdx.x & 31; }

}  // namespace b3b200
