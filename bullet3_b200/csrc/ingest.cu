// ingest.cu -- the callers' data formats either side of the step:
//   * b3b200_register_concave_obj   Wavefront .obj -> trimesh collidable, the way ConcaveScene::createConcaveMesh feeds
//                                   registerConcaveMesh (examples/OpenCL/rigidbody/ConcaveScene.cpp:28-109, 111-158):
//                                   triangle soup (three fresh vertices per face corner), vertex = (p + shift) * scaling
//   * b3b200_checkpoint_save/load   body buffer (+ inertias + joints) dump / restore; the reference has no counterpart for
//                                   the GPU pipeline (its .bullet serializer belongs to Bullet 2)
//   * b3b200_copy_transforms        copyTransformsToVBOKernel (examples/OpenCL/rigidbody/GpuRigidBodyDemo.cpp:52-60):
//                                   instance positions and orientations into a caller-owned DEVICE buffer (a mapped
//                                   graphics-interop buffer or any other device allocation)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "internal.h"
#include "../../include/b3b200.h"

using namespace b3b200;

namespace
{
constexpr int COPY_THREADS = 256;

__global__ void __launch_bounds__(COPY_THREADS) copyTransformsKernel(const float4* __restrict__ pose, float4* __restrict__ posOrnColor, int numNodes)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= numNodes) return;
	const float4 p = pose[2 * i];
	posOrnColor[i] = make_float4(p.x, p.y, p.z, 1.0f);
	posOrnColor[i + numNodes] = pose[2 * i + 1];
}

struct CheckpointHeader
{
	char magic[8];
	int version, numBodies, numJoints, jointUid;
	int bodyBytes, inertiaBytes, jointBytes, reserved;
};
const char MAGIC[8] = {'B', '3', 'B', '2', '0', '0', 'C', 'P'};

// one corner of an `f` record: v, v/vt, v//vn or v/vt/vn; 1-based, negative = relative to the vertices read so far
bool parseCorner(const char*& s, int numPositions, int& index)
{
	while (*s == ' ' || *s == '\t') s++;
	if (!*s || *s == '\n' || *s == '\r' || *s == '#') return false;
	char* end = 0;
	long v = strtol(s, &end, 10);
	if (end == s) return false;
	s = end;
	while (*s && *s != ' ' && *s != '\t' && *s != '\n' && *s != '\r') s++;  // skip /vt/vn
	index = v > 0 ? (int)v - 1 : numPositions + (int)v;
	return true;
}
}  // namespace

extern "C" int b3b200_register_concave_obj(b3b200_world* w, const char* path, const float* shift3, const float* scaling3)
{
	if (!w || !path) return -1;
	FILE* f = fopen(path, "r");
	if (!f)
	{
		setLastError("register_concave_obj: cannot open file");
		return -1;
	}
	const float sh[3] = {shift3 ? shift3[0] : 0.f, shift3 ? shift3[1] : 0.f, shift3 ? shift3[2] : 0.f};
	const float sc[3] = {scaling3 ? scaling3[0] : 1.f, scaling3 ? scaling3[1] : 1.f, scaling3 ? scaling3[2] : 1.f};
	std::vector<float> positions, soup;
	std::vector<int> indices, corners;
	std::vector<char> line(1 << 16);
	bool bad = false;
	while (fgets(line.data(), (int)line.size(), f))
	{
		const char* s = line.data();
		while (*s == ' ' || *s == '\t') s++;
		if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t'))
		{
			float x = 0, y = 0, z = 0;
			if (sscanf(s + 1, "%f %f %f", &x, &y, &z) != 3)
			{
				bad = true;
				break;
			}
			positions.push_back(x);
			positions.push_back(y);
			positions.push_back(z);
		}
		else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t'))
		{
			s++;
			corners.clear();
			int idx;
			const int np = (int)positions.size() / 3;
			while (parseCorner(s, np, idx))
			{
				if (idx < 0 || idx >= np)
				{
					bad = true;
					break;
				}
				corners.push_back(idx);
			}
			if (bad) break;
			// polygons become fans around their first corner (tinyobj's triangulation)
			for (size_t k = 2; k < corners.size(); k++)
			{
				const int tri[3] = {corners[0], corners[k - 1], corners[k]};
				for (int c = 0; c < 3; c++)
				{
					indices.push_back((int)soup.size() / 3);
					for (int j = 0; j < 3; j++) soup.push_back((positions[3 * tri[c] + j] + sh[j]) * sc[j]);
				}
			}
		}
	}
	fclose(f);
	if (bad || indices.empty())
	{
		setLastError(bad ? "register_concave_obj: malformed v / f record" : "register_concave_obj: no faces in file");
		return -1;
	}
	const float one[3] = {1.f, 1.f, 1.f};
	return b3b200_register_concave(w, soup.data(), (int)soup.size() / 3, indices.data(), (int)indices.size(), one);
}

extern "C" int b3b200_checkpoint_save(b3b200_world* w, const char* path)
{
	if (!w || !path) return B3B200_ERR_INVALID;
	if (w->device < 0 || !w->uploaded) return B3B200_ERR_STATE;
	const int n = w->numBodies;
	std::vector<b3b200_rigid_body> bodies((size_t)n);
	std::vector<b3b200_inertia> inertias((size_t)n);
	B3_TRY(b3b200_readback_bodies(w, bodies.data(), n));
	B3_TRY(b3b200_readback_inertias(w, inertias.data(), n));
	int numJoints = 0;
	B3_TRY(b3b200_get_joints(w, 0, 0, &numJoints));  // refreshes the host copy (flags of broken joints)
	CheckpointHeader h;
	memset(&h, 0, sizeof(h));
	memcpy(h.magic, MAGIC, 8);
	h.version = 1;
	h.numBodies = n;
	h.numJoints = numJoints;
	h.jointUid = w->jointUid;
	h.bodyBytes = (int)sizeof(b3b200_rigid_body);
	h.inertiaBytes = (int)sizeof(b3b200_inertia);
	h.jointBytes = (int)sizeof(b3b200_generic_constraint);
	FILE* f = fopen(path, "wb");
	if (!f)
	{
		setLastError("checkpoint_save: cannot open file");
		return B3B200_ERR_INVALID;
	}
	bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
	ok = ok && (n == 0 || fwrite(bodies.data(), sizeof(b3b200_rigid_body), (size_t)n, f) == (size_t)n);
	ok = ok && (n == 0 || fwrite(inertias.data(), sizeof(b3b200_inertia), (size_t)n, f) == (size_t)n);
	ok = ok && (numJoints == 0 || fwrite(w->joints.data(), sizeof(b3b200_generic_constraint), (size_t)numJoints, f) == (size_t)numJoints);
	ok = (fclose(f) == 0) && ok;
	if (!ok)
	{
		setLastError("checkpoint_save: short write");
		return B3B200_ERR_INVALID;
	}
	return 0;
}

// The world must hold the same shapes and the same number of bodies as the one that was saved (shapes are set-up data,
// re-registered by the caller); body state, inertias and the joint set are replaced.
extern "C" int b3b200_checkpoint_load(b3b200_world* w, const char* path)
{
	if (w) w->dropStepGraphs();  // may change what a step launches
	if (!w || !path) return B3B200_ERR_INVALID;
	if (w->device < 0 || !w->uploaded) return B3B200_ERR_STATE;
	FILE* f = fopen(path, "rb");
	if (!f)
	{
		setLastError("checkpoint_load: cannot open file");
		return B3B200_ERR_INVALID;
	}
	CheckpointHeader h;
	bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, MAGIC, 8) == 0 && h.version == 1;
	ok = ok && h.bodyBytes == (int)sizeof(b3b200_rigid_body) && h.inertiaBytes == (int)sizeof(b3b200_inertia) && h.jointBytes == (int)sizeof(b3b200_generic_constraint);
	if (!ok || h.numBodies != w->numBodies || h.numJoints < 0)
	{
		fclose(f);
		setLastError(!ok ? "checkpoint_load: not a checkpoint of this build" : "checkpoint_load: body count differs from the world's");
		return B3B200_ERR_INVALID;
	}
	const size_t n = (size_t)h.numBodies, nj = (size_t)h.numJoints;
	std::vector<b3b200_rigid_body> bodies(n);
	std::vector<b3b200_inertia> inertias(n);
	std::vector<b3b200_generic_constraint> joints(nj);
	ok = (n == 0 || fread(bodies.data(), sizeof(b3b200_rigid_body), n, f) == n) && (n == 0 || fread(inertias.data(), sizeof(b3b200_inertia), n, f) == n) &&
		 (nj == 0 || fread(joints.data(), sizeof(b3b200_generic_constraint), nj, f) == nj);
	fclose(f);
	if (!ok)
	{
		setLastError("checkpoint_load: truncated file");
		return B3B200_ERR_INVALID;
	}
	for (size_t i = 0; i < n; i++)
		if (bodies[i].collidableIdx < 0 || bodies[i].collidableIdx >= (int)w->collidables.size())
		{
			setLastError("checkpoint_load: a body refers to a collidable this world does not have");
			return B3B200_ERR_INVALID;
		}
	for (size_t j = 0; j < nj; j++)
		if (joints[j].rbA < 0 || joints[j].rbA >= (int)n || joints[j].rbB < 0 || joints[j].rbB >= (int)n)
		{
			setLastError("checkpoint_load: a joint refers to a body this world does not have");
			return B3B200_ERR_INVALID;
		}
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	if (n)
	{
		B3_TRY(b3b200_write_bodies(w, bodies.data(), (int)n));
		B3_CUDA_CHECK(cudaMemcpyAsync(w->dInertias.ptr, inertias.data(), sizeof(b3b200_inertia) * n, cudaMemcpyHostToDevice, w->stream));
		B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
		w->inertias = inertias;
	}
	w->joints = joints;
	w->jointUid = h.jointUid;
	w->jointsDirty = true;
	w->jointBatchesDirty = true;
	return 0;
}

extern "C" int b3b200_copy_transforms(b3b200_world* w, void* dstDevice, int numNodes)
{
	if (!w || numNodes < 0 || (numNodes && !dstDevice)) return B3B200_ERR_INVALID;
	if (w->device < 0 || !w->uploaded) return B3B200_ERR_STATE;
	if (numNodes > w->numBodies) return B3B200_ERR_INVALID;
	if (numNodes == 0) return 0;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	copyTransformsKernel<<<divUp(numNodes, COPY_THREADS), COPY_THREADS, 0, w->stream>>>(w->dPose.ptr, (float4*)dstDevice, numNodes);
	B3_LAUNCH_CHECK();
	return 0;
}
