// jacobi.cu -- mass-splitting Jacobi contact solver (Tonge 2012).
//
// Replaces b3GpuJacobiContactSolver::solveContacts
// (src/Bullet3OpenCL/RigidBody/b3GpuJacobiContactSolver.cpp:699-869) and its kernels in
// kernels/solverUtils.cl: CountBodiesKernel :392, ContactToConstraintSplitKernel :926
// (setConstraint4 with contact counts :836-923), ClearVelocitiesKernel :416,
// SolveContactJacobiKernel :641, AverageVelocitiesKernel :428, SolveFrictionJacobiKernel :801,
// UpdateBodyVelocitiesKernel :815.  Same order per iteration: contacts, average, friction, average.
// Constants of b3JacobiSolverInfo (b3GpuJacobiContactSolver.h:16-33): dt 1/60, drift 0.005,
// positionConstraintCoeff 0.99; the iteration count is the world's setting (reference default 7).
// No colouring is needed: every contact works on its own split copy of the body's delta velocity.
#include "internal.h"

namespace b3b200
{
constexpr int JT = 256;

B3_D float4 jMatMul(const float4* I, const float4& v) { return mk4(dot3(I[0], v), dot3(I[1], v), dot3(I[2], v)); }
B3_D float jRelVel(const float4& l0, const float4& l1, const float4& a0, const float4& a1, const float4& lv0, const float4& av0, const float4& lv1, const float4& av1)
{
	return dot3(l0, lv0) + dot3(a0, av0) + dot3(l1, lv1) + dot3(a1, av1);
}
B3_D void jPlaneSpace1(const float4& n, float4& p, float4& q)
{
	if (fabsf(n.z) > 0.70710678f)
	{
		float a = n.y * n.y + n.z * n.z;
		float k = 1.f / sqrtf(a);
		p = mk4(0.f, -n.z * k, n.y * k);
		q = mk4(a * k, -n.x * p.z, n.x * p.y);
	}
	else
	{
		float a = n.x * n.x + n.y * n.y;
		float k = 1.f / sqrtf(a);
		p = mk4(-n.y * k, n.x * k, 0.f);
		q = mk4(-n.z * p.y, n.z * p.x, a * k);
	}
}
// calcJacCoeff with mass splitting (solverUtils.cl:474-483)
B3_D float jJacCoeff(const float4& angular0, const float4& angular1, float invMass0, const float4* I0, float invMass1, const float4* I1, float countA, float countB)
{
	float jmj0 = invMass0;
	float jmj1 = dot3(jMatMul(I0, angular0), angular0);
	float jmj2 = invMass1;
	float jmj3 = dot3(jMatMul(I1, angular1), angular1);
	return -1.f / ((jmj0 + jmj1) * countA + (jmj2 + jmj3) * countB);
}

struct JArgs
{
	const b3b200_contact4* contacts;
	const unsigned int* ctr;
	const float4* pose;
	float4* vel;
	const b3b200_inertia* inertias;
	b3b200_constraint4* constraints;
	unsigned int* bodyCount;
	const unsigned int* bodyOffset;
	int2* slots;
	float4* dLin;
	float4* dAng;
	int numBodies;
	int staticIdx;
	float dt, positionDrift, positionConstraintCoeff;
};

__global__ void __launch_bounds__(JT) jCountKernel(JArgs s)
{
	const int n = (int)s.ctr[CTR_CONTACTS];
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const int4 ids = reinterpret_cast<const int4*>(&s.contacts[i])[5];
		int2 slot = make_int2(0, 0);
		if (!(ids.z < 0 || ids.z == s.staticIdx)) slot.x = (int)atomicAdd(&s.bodyCount[abs(ids.z)], 1u);
		if (!(ids.w < 0 || ids.w == s.staticIdx)) slot.y = (int)atomicAdd(&s.bodyCount[abs(ids.w)], 1u);
		s.slots[i] = slot;
	}
}

// setConstraint4 + ContactToConstraintSplitKernel (solverUtils.cl:836-967)
__global__ void __launch_bounds__(JT) jBuildKernel(JArgs s)
{
	const int n = (int)s.ctr[CTR_CONTACTS];
	for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x)
	{
		const float4* cw = reinterpret_cast<const float4*>(&s.contacts[g]);
		const float4 wp[4] = {cw[0], cw[1], cw[2], cw[3]};
		const float4 nrm = cw[4];
		const int4 ids = reinterpret_cast<const int4*>(&s.contacts[g])[5];
		const int aIdx = abs(ids.z), bIdx = abs(ids.w);
		const float4 posA = s.pose[2 * aIdx], posB = s.pose[2 * bIdx];
		const float4 linVelA = s.vel[2 * aIdx], angVelA = s.vel[2 * aIdx + 1];
		const float4 linVelB = s.vel[2 * bIdx], angVelB = s.vel[2 * bIdx + 1];
		const float invMassA = posA.w, invMassB = posB.w;
		const float4* IAp = reinterpret_cast<const float4*>(&s.inertias[aIdx].invInertiaWorld);
		const float4* IBp = reinterpret_cast<const float4*>(&s.inertias[bIdx].invInertiaWorld);
		const float4 IA[3] = {__ldg(IAp), __ldg(IAp + 1), __ldg(IAp + 2)};
		const float4 IB[3] = {__ldg(IBp), __ldg(IBp + 1), __ldg(IBp + 2)};
		const float countA = invMassA != 0.f ? (float)s.bodyCount[aIdx] : 1.f;
		const float countB = invMassB != 0.f ? (float)s.bodyCount[bIdx] : 1.f;
		const float dtInv = 1.f / s.dt;
		const float npoints = nrm.w;
		const float4 n3 = mk4(nrm.x, nrm.y, nrm.z);
		float jac[4], bb[4];
#pragma unroll
		for (int ic = 0; ic < 4; ic++)
		{
			const float4 r0 = sub3(wp[ic], posA), r1 = sub3(wp[ic], posB);
			if ((float)ic >= npoints)
			{
				jac[ic] = 0.f;
				bb[ic] = 0.f;
				continue;
			}
			const float4 a0 = cross3(r0, n3), a1 = neg3(cross3(r1, n3));
			jac[ic] = jJacCoeff(a0, a1, invMassA, IA, invMassB, IB, countA, countB);
			const float relVelN = jRelVel(n3, neg3(n3), a0, a1, linVelA, angVelA, linVelB, angVelB);
			const float e = 0.f;
			bb[ic] = e * relVelN + (wp[ic].w + s.positionDrift) * s.positionConstraintCoeff * dtInv;
		}
		float fjac[2] = {0.f, 0.f};
		float4 center = mk4(0, 0, 0);
		if (npoints > 0)
		{
			for (int i = 0; (float)i < npoints && i < 4; i++) center = add3(center, wp[i]);
			const float inv = 1.0f / npoints;
			center = scale3(center, inv);
			float4 t0, t1;
			jPlaneSpace1(neg3(n3), t0, t1);
			const float4 r0 = sub3(center, posA), r1 = sub3(center, posB);
			fjac[0] = jJacCoeff(cross3(r0, t0), neg3(cross3(r1, t0)), invMassA, IA, invMassB, IB, countA, countB);
			fjac[1] = jJacCoeff(cross3(r0, t1), neg3(cross3(r1, t1)), invMassA, IA, invMassB, IB, countA, countB);
		}
		float4* dw = reinterpret_cast<float4*>(&s.constraints[g]);
		dw[0] = mk4(nrm.x, nrm.y, nrm.z, 0.7f);
#pragma unroll
		for (int i = 0; i < 4; i++) dw[1 + i] = ((float)i < npoints) ? wp[i] : mk4(0, 0, 0, 0);
		dw[5] = center;
		dw[6] = mk4(jac[0], jac[1], jac[2], jac[3]);
		dw[7] = mk4(bb[0], bb[1], bb[2], bb[3]);
		dw[8] = mk4(0, 0, 0, 0);
		dw[9] = mk4(fjac[0], fjac[1], 0, 0);
		int4 tail;
		tail.x = aIdx;
		tail.y = bIdx;
		tail.z = reinterpret_cast<const int4*>(&s.contacts[g])[5].y;  // batchIdx
		tail.w = 0;
		reinterpret_cast<int4*>(dw)[10] = tail;
	}
}

// SolveContactJacobiKernel / SolveFrictionJacobiKernel (solverUtils.cl:527-811)
template <bool FRICTION>
__global__ void __launch_bounds__(JT) jSolveKernel(JArgs s)
{
	const int n = (int)s.ctr[CTR_CONTACTS];
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		float4* cw = reinterpret_cast<float4*>(&s.constraints[i]);
		const int4 tail = reinterpret_cast<const int4*>(cw)[10];
		const int aIdx = tail.x, bIdx = tail.y;
		const float4 posA = s.pose[2 * aIdx], posB = s.pose[2 * bIdx];
		const float invMassA = posA.w, invMassB = posB.w;
		const float4 linVelA = s.vel[2 * aIdx], angVelA = s.vel[2 * aIdx + 1];
		const float4 linVelB = s.vel[2 * bIdx], angVelB = s.vel[2 * bIdx + 1];
		const float4* IAp = reinterpret_cast<const float4*>(&s.inertias[aIdx].invInertiaWorld);
		const float4* IBp = reinterpret_cast<const float4*>(&s.inertias[bIdx].invInertiaWorld);
		const float4 IA[3] = {__ldg(IAp), __ldg(IAp + 1), __ldg(IAp + 2)};
		const float4 IB[3] = {__ldg(IBp), __ldg(IBp + 1), __ldg(IBp + 2)};
		const int2 slot = s.slots[i];
		const int splitA = (int)s.bodyOffset[aIdx] + slot.x, splitB = (int)s.bodyOffset[bIdx] + slot.y;
		float4 dLinA = mk4(0, 0, 0), dAngA = mk4(0, 0, 0), dLinB = mk4(0, 0, 0), dAngB = mk4(0, 0, 0);
		if (invMassA != 0.f)
		{
			dLinA = s.dLin[splitA];
			dAngA = s.dAng[splitA];
		}
		if (invMassB != 0.f)
		{
			dLinB = s.dLin[splitB];
			dAngB = s.dAng[splitB];
		}
		const float4 lin = cw[0];
		const float4 n3 = mk4(lin.x, lin.y, lin.z);
		if (!FRICTION)
		{
			const float4 jac = cw[6], bias = cw[7];
			float4 applied = cw[8];
			const float jacv[4] = {jac.x, jac.y, jac.z, jac.w};
			const float bv[4] = {bias.x, bias.y, bias.z, bias.w};
			float ap[4] = {applied.x, applied.y, applied.z, applied.w};
#pragma unroll
			for (int ic = 0; ic < 4; ic++)
			{
				if (jacv[ic] == 0.f) continue;
				const float4 wp = cw[1 + ic];
				const float4 r0 = sub3(wp, posA), r1 = sub3(wp, posB);
				const float4 a0 = cross3(r0, n3), a1 = neg3(cross3(r1, n3));
				float rambdaDt = jRelVel(n3, neg3(n3), a0, a1, add3(linVelA, dLinA), add3(angVelA, dAngA), add3(linVelB, dLinB), add3(angVelB, dAngB)) + bv[ic];
				rambdaDt *= jacv[ic];
				const float prevSum = ap[ic];
				float updated = prevSum + rambdaDt;
				updated = fmaxf(updated, 0.f);
				updated = fminf(updated, FLT_MAX);
				rambdaDt = updated - prevSum;
				ap[ic] = updated;
				if (invMassA != 0.f)
				{
					dLinA = add3(dLinA, scale3(scale3(n3, invMassA), rambdaDt));
					dAngA = add3(dAngA, scale3(jMatMul(IA, a0), rambdaDt));
				}
				if (invMassB != 0.f)
				{
					dLinB = add3(dLinB, scale3(scale3(neg3(n3), invMassB), rambdaDt));
					dAngB = add3(dAngB, scale3(jMatMul(IB, a1), rambdaDt));
				}
			}
			cw[8] = mk4(ap[0], ap[1], ap[2], ap[3]);
		}
		else
		{
			float4 fr = cw[9];
			if (!(fr.x == 0.f && fr.x == 0.f))
			{
				const float4 applied = cw[8];
				float sum = 0.f;
				sum += applied.x;
				sum += applied.y;
				sum += applied.z;
				sum += applied.w;
				const float maxR = 0.7f * sum, minR = -maxR;
				const float4 center = cw[5];
				const float4 nn = neg3(n3);
				float4 t[2];
				jPlaneSpace1(nn, t[0], t[1]);
				const float4 r0 = sub3(center, posA), r1 = sub3(center, posB);
				float fj[2] = {fr.x, fr.y}, fa[2] = {fr.z, fr.w};
#pragma unroll
				for (int k = 0; k < 2; k++)
				{
					const float4 a0 = cross3(r0, t[k]), a1 = neg3(cross3(r1, t[k]));
					float rambdaDt = jRelVel(t[k], neg3(t[k]), a0, a1, add3(linVelA, dLinA), add3(angVelA, dAngA), add3(linVelB, dLinB), add3(angVelB, dAngB));
					rambdaDt *= fj[k];
					const float prevSum = fa[k];
					float updated = prevSum + rambdaDt;
					updated = fmaxf(updated, minR);
					updated = fminf(updated, maxR);
					rambdaDt = updated - prevSum;
					fa[k] = updated;
					dLinA = add3(dLinA, scale3(scale3(t[k], invMassA), rambdaDt));
					dLinB = add3(dLinB, scale3(scale3(neg3(t[k]), invMassB), rambdaDt));
					dAngA = add3(dAngA, scale3(jMatMul(IA, a0), rambdaDt));
					dAngB = add3(dAngB, scale3(jMatMul(IB, a1), rambdaDt));
				}
				{
					const float4 ab = normalized3(sub3(posB, posA));
					const float4 ac = normalized3(sub3(center, posA));
					if (dot3(ab, ac) > 0.95f || (invMassA == 0.f || invMassB == 0.f))
					{
						const float angNA = dot3(nn, angVelA), angNB = dot3(nn, angVelB);
						dAngA = sub3(dAngA, scale3(nn, angNA * 0.1f));
						dAngB = sub3(dAngB, scale3(nn, angNB * 0.1f));
					}
				}
				cw[9] = mk4(fj[0], fj[1], fa[0], fa[1]);
			}
			else
				continue;  // the reference returns before writing the deltas back
		}
		if (invMassA != 0.f)
		{
			s.dLin[splitA] = dLinA;
			s.dAng[splitA] = dAngA;
		}
		if (invMassB != 0.f)
		{
			s.dLin[splitB] = dLinB;
			s.dAng[splitB] = dAngB;
		}
	}
}

// AverageVelocitiesKernel (solverUtils.cl:428-456); FINAL = UpdateBodyVelocitiesKernel (:815-833) fused after the last average
template <bool FINAL>
__global__ void __launch_bounds__(JT) jAverageKernel(JArgs s)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s.numBodies; i += gridDim.x * blockDim.x)
	{
		if (s.pose[2 * i].w == 0.f) continue;
		const int off = (int)s.bodyOffset[i];
		const int count = (int)s.bodyCount[i];
		if (count == 0) continue;
		const float factor = 1.f / ((float)count);
		float4 avgL = mk4(0, 0, 0), avgA = mk4(0, 0, 0);
		for (int j = 0; j < count; j++)
		{
			avgL = add3(avgL, scale3(s.dLin[off + j], factor));
			avgA = add3(avgA, scale3(s.dAng[off + j], factor));
		}
		if (FINAL)
		{
			s.vel[2 * i] = add3(s.vel[2 * i], avgL);
			s.vel[2 * i + 1] = add3(s.vel[2 * i + 1], avgA);
		}
		else
			for (int j = 0; j < count; j++)
			{
				s.dLin[off + j] = avgL;
				s.dAng[off + j] = avgA;
			}
	}
}

int launchJacobi(World* w)
{
	cudaStream_t st = w->stream;
	const int nb = std::max(w->numBodies, 1);
	JArgs s;
	s.contacts = w->dContacts.ptr;
	s.ctr = w->dCounters.ptr;
	s.pose = w->dPose.ptr;
	s.vel = w->dVel.ptr;
	s.inertias = w->dInertias.ptr;
	s.constraints = w->dConstraints.ptr;
	s.bodyCount = w->dBodyCount.ptr;
	s.bodyOffset = w->dBodyOffset.ptr;
	s.slots = reinterpret_cast<int2*>(w->dContactSlot.ptr);
	s.dLin = w->dDeltaLin.ptr;
	s.dAng = w->dDeltaAng.ptr;
	s.numBodies = w->numBodies;
	s.staticIdx = w->static0Index;
	s.dt = 1.f / 60.f;
	s.positionDrift = 0.005f;
	s.positionConstraintCoeff = 0.99f;
	const int grid = w->smCount * 8;
	const size_t splitCap = 2 * (size_t)std::max(w->cfg.maxContactCapacity, 1);
	// Jacobi-only buffers are allocated on first use (0.27 GB at the default 16 contacts/body capacity)
	B3_TRY(w->dBodyOffset.reserve(nb));
	B3_TRY(w->dConstraints.reserve((size_t)std::max(w->cfg.maxContactCapacity, 1)));
	s.constraints = w->dConstraints.ptr;
	B3_TRY(w->dContactSlot.reserve(splitCap));
	B3_TRY(w->dDeltaLin.reserve(splitCap));
	B3_TRY(w->dDeltaAng.reserve(splitCap));
	s.bodyOffset = w->dBodyOffset.ptr;
	s.slots = reinterpret_cast<int2*>(w->dContactSlot.ptr);
	s.dLin = w->dDeltaLin.ptr;
	s.dAng = w->dDeltaAng.ptr;
	B3_CUDA_CHECK(cudaMemsetAsync(w->dBodyCount.ptr, 0, sizeof(unsigned int) * nb, st));
	jCountKernel<<<grid, JT, 0, st>>>(s);
	B3_LAUNCH_CHECK();
	B3_TRY(exclusiveScanU32(st, w->dBodyCount.ptr, w->dBodyOffset.ptr, w->numBodies, nullptr));
	jBuildKernel<<<grid, JT, 0, st>>>(s);
	B3_LAUNCH_CHECK();
	B3_CUDA_CHECK(cudaMemsetAsync(w->dDeltaLin.ptr, 0, sizeof(float4) * splitCap, st));
	B3_CUDA_CHECK(cudaMemsetAsync(w->dDeltaAng.ptr, 0, sizeof(float4) * splitCap, st));
	for (int iter = 0; iter < w->solverIterations; iter++)
	{
		jSolveKernel<false><<<grid, JT, 0, st>>>(s);
		B3_LAUNCH_CHECK();
		jAverageKernel<false><<<grid, JT, 0, st>>>(s);
		B3_LAUNCH_CHECK();
		jSolveKernel<true><<<grid, JT, 0, st>>>(s);
		B3_LAUNCH_CHECK();
		if (iter == w->solverIterations - 1)
			jAverageKernel<true><<<grid, JT, 0, st>>>(s);
		else
			jAverageKernel<false><<<grid, JT, 0, st>>>(s);
		B3_LAUNCH_CHECK();
	}
	w->soaDirty = true;
	return 0;
}

}  // namespace b3b200
