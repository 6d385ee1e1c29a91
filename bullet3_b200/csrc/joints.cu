// joints.cu -- point-to-point and fixed joints: b3GpuPgsConstraintSolver::solveJoints
// (src/Bullet3OpenCL/RigidBody/b3GpuPgsConstraintSolver.cpp:927-942 -> solveGroup :909-925 ->
// solveGroupCacheFriendlySetup :191-565, ...Iterations :616-776, ...Finish :944-1068, kernels/jointSolver.cl)
// and the constraint bookkeeping of b3GpuRigidBodyPipeline (b3GpuRigidBodyPipeline.cpp:158-218).
//
// The reference runs 7 kernels + a host-side batching pass with a clFinish after each; here
//   * the batches are built on the host when the constraint set changes, with the reference's own greedy rule
//     (sortConstraintByBatch3 :778-905: walk the constraints in order, a constraint joins the current batch when none
//     of its dynamic bodies is used by it yet) -- so the solve order, and with it the result, is the reference's;
//   * jointSetupKernel   one thread per constraint: getInfo2 rows (P2P: 3, fixed: 6) + row finalisation
//   * jointSolveKernel   ONE CTA runs all iterations x batches with __syncthreads() in between when the largest batch
//                        fits (the usual case: joints are few), else one launch per batch
//   * jointFinishKernel  break violated joints, v += dv
// Solver-body velocity deltas live in a 32-byte-per-body array (zeroed per step, only when joints exist).
#include <string.h>
#include "internal.h"

namespace b3b200
{
namespace
{
constexpr int J_THREADS = 256;
constexpr int J_MAX_ROWS = 6;
#define B3B200_INFINITY FLT_MAX  // B3_INFINITY (Bullet3Common/b3Scalar.h)

// row of b3GpuSolverConstraint (b3GpuSolverConstraint.h:26-71) reduced to what the joint path uses
struct JointRow
{
	float4 normal;       // m_contactNormal        (w = m_jacDiagABInv)
	float4 rel1;         // m_relpos1CrossNormal   (w = m_rhs)
	float4 rel2;         // m_relpos2CrossNormal   (w = m_cfm)
	float4 angA;         // m_angularComponentA    (w = m_lowerLimit)
	float4 angB;         // m_angularComponentB    (w = m_upperLimit)
	float appliedImpulse;
	int bodyA, bodyB, pad;
};
static_assert(sizeof(JointRow) == 96, "row layout");

struct JointArgs
{
	b3b200_generic_constraint* joints;
	int numJoints;
	const int* order;        // constraint indices in batch order
	const int* batchOffset;  // numBatches + 1
	int numBatches;
	JointRow* rows;          // J_MAX_ROWS per constraint
	int* numRows;            // per constraint (info1)
	float4* delta;           // 2 per body: delta linear / angular velocity
	const float4* pose;
	float4* vel;
	const b3b200_inertia* inertias;
	int numBodies;
	int iterations;
	float fps, erp, cfm, damping;
};

B3_D float4 matVec(const float4* I, const float4& v) { return mk4(dot3(I[0], v), dot3(I[1], v), dot3(I[2], v)); }

// getInfo1 + getInfo2 + "finalize the constraint setup" (b3GpuPgsConstraintSolver.cpp:296-541, jointSolver.cl:374-398, 488-543,
// 592-628, 651-877), one constraint
B3_D void setupOne(const JointArgs& a, int ci)
{
	const b3b200_generic_constraint c = a.joints[ci];
	JointRow* rows = a.rows + (size_t)ci * J_MAX_ROWS;
	int n = 0;
	if (c.flags & B3B200_CONSTRAINT_FLAG_ENABLED) n = c.constraintType == B3B200_CONSTRAINT_P2P ? 3 : (c.constraintType == B3B200_CONSTRAINT_FIXED ? 6 : 0);
	a.numRows[ci] = n;
	if (!n) return;
	const float4 posA = a.pose[2 * c.rbA], ornA = a.pose[2 * c.rbA + 1];
	const float4 posB = a.pose[2 * c.rbB], ornB = a.pose[2 * c.rbB + 1];
	const float invMassA = posA.w, invMassB = posB.w;
	const float4 linVelA = a.vel[2 * c.rbA], angVelA = a.vel[2 * c.rbA + 1];
	const float4 linVelB = a.vel[2 * c.rbB], angVelB = a.vel[2 * c.rbB + 1];
	const float4* IA = reinterpret_cast<const float4*>(&a.inertias[c.rbA].invInertiaWorld);
	const float4* IB = reinterpret_cast<const float4*>(&a.inertias[c.rbB].invInertiaWorld);
	const float k = a.fps * a.erp;

	// getInfo2Point2Point: rows 0..2
	const Mat3 mA = matFromQuat(ornA), mB = matFromQuat(ornB);
	const float4 a1 = matMulVec(mA, mk4(c.pivotInA.x, c.pivotInA.y, c.pivotInA.z));
	const float4 a2 = matMulVec(mB, mk4(c.pivotInB.x, c.pivotInB.y, c.pivotInB.z));
	const float4 a1n = neg3(a1);
	// b3Vector3::getSkewSymmetricMatrix (b3Vector3.h:633-638): v0 = (0,-z,y), v1 = (z,0,-x), v2 = (-y,x,0)
	float4 J1a[3] = {mk4(0.f, -a1n.z, a1n.y), mk4(a1n.z, 0.f, -a1n.x), mk4(-a1n.y, a1n.x, 0.f)};
	float4 J2a[3] = {mk4(0.f, -a2.z, a2.y), mk4(a2.z, 0.f, -a2.x), mk4(-a2.y, a2.x, 0.f)};
	const float err[3] = {k * (a2.x + posB.x - a1.x - posA.x), k * (a2.y + posB.y - a1.y - posA.y), k * (a2.z + posB.z - a1.z - posA.z)};
	float4 normals[J_MAX_ROWS], rel1[J_MAX_ROWS], rel2[J_MAX_ROWS];
	float rhs[J_MAX_ROWS];
	for (int j = 0; j < 3; j++)
	{
		normals[j] = mk4(j == 0 ? 1.f : 0.f, j == 1 ? 1.f : 0.f, j == 2 ? 1.f : 0.f);
		rel1[j] = J1a[j];
		rel2[j] = J2a[j];
		rhs[j] = err[j];
	}
	if (n == 6)
	{
		// getInfo2FixedOrientation (jointSolver.cl:592-628): 3 rows that make the body rotations equal
		const float4 qrelCur = quatMul(ornA, quatInverse(ornB));
		const float4 q0 = mk4(c.relTargetAB.x, c.relTargetAB.y, c.relTargetAB.z, c.relTargetAB.w);
		// nearest(): qd or -qd, whichever is closer to q0
		const float4 dq = mk4(q0.x - qrelCur.x, q0.y - qrelCur.y, q0.z - qrelCur.z, q0.w - qrelCur.w);
		const float4 sq = mk4(q0.x + qrelCur.x, q0.y + qrelCur.y, q0.z + qrelCur.z, q0.w + qrelCur.w);
		const float dd = dq.x * dq.x + dq.y * dq.y + dq.z * dq.z + dq.w * dq.w, ss = sq.x * sq.x + sq.y * sq.y + sq.z * sq.z + sq.w * sq.w;
		const float4 orn1 = dd < ss ? qrelCur : mk4(-qrelCur.x, -qrelCur.y, -qrelCur.z, -qrelCur.w);
		float4 dorn = quatMul(orn1, quatInverse(q0));
		if (dorn.w >= 1.f) dorn.w = 1.f;
		float x = dorn.w;
		if (x < -1.f) x = -1.f;
		if (x > 1.f) x = 1.f;
		const float angle = 2.f * acosf(x);
		float4 axis = mk4(dorn.x, dorn.y, dorn.z);
		const float len = dot3(axis, axis);
		if (len < FLT_EPSILON * FLT_EPSILON)
			axis = mk4(1.f, 0.f, 0.f);
		else
		{
			const float sl = sqrtf(len);
			axis = mk4(axis.x / sl, axis.y / sl, axis.z / sl);
		}
		const float4 diff = scale3(axis, -angle);
		const float e3[3] = {k * diff.x, k * diff.y, k * diff.z};
		for (int j = 0; j < 3; j++)
		{
			normals[3 + j] = mk4(0, 0, 0);
			rel1[3 + j] = mk4(j == 0 ? 1.f : 0.f, j == 1 ? 1.f : 0.f, j == 2 ? 1.f : 0.f);
			rel2[3 + j] = mk4(j == 0 ? -1.f : 0.f, j == 1 ? -1.f : 0.f, j == 2 ? -1.f : 0.f);
			rhs[3 + j] = e3[j];
		}
	}
	for (int j = 0; j < n; j++)
	{
		float lower = -B3B200_INFINITY, upper = B3B200_INFINITY;
		if (upper >= c.breakingImpulseThreshold) upper = c.breakingImpulseThreshold;
		if (lower <= -c.breakingImpulseThreshold) lower = -c.breakingImpulseThreshold;
		const float4 angCompA = matVec(IA, rel1[j]);
		const float4 angCompB = matVec(IB, rel2[j]);
		const float4 iMJlA = scale3(normals[j], invMassA);
		const float4 iMJlB = scale3(normals[j], invMassB);
		float sum = dot3(iMJlA, normals[j]);
		sum += dot3(angCompA, rel1[j]);
		sum += dot3(iMJlB, normals[j]);
		sum += dot3(angCompB, rel2[j]);
		const float jacDiagABInv = fabsf(sum) > FLT_EPSILON ? 1.f / sum : 0.f;
		const float vel1Dotn = dot3(normals[j], linVelA) + dot3(rel1[j], angVelA);
		const float vel2Dotn = -dot3(normals[j], linVelB) + dot3(rel2[j], angVelB);
		const float relVel = vel1Dotn + vel2Dotn;
		const float velocityError = 0.f - relVel * a.damping;
		const float penetrationImpulse = rhs[j] * jacDiagABInv;
		const float velocityImpulse = velocityError * jacDiagABInv;
		JointRow r;
		r.normal = mk4(normals[j].x, normals[j].y, normals[j].z, jacDiagABInv);
		r.rel1 = mk4(rel1[j].x, rel1[j].y, rel1[j].z, penetrationImpulse + velocityImpulse);
		r.rel2 = mk4(rel2[j].x, rel2[j].y, rel2[j].z, a.cfm);
		r.angA = mk4(angCompA.x, angCompA.y, angCompA.z, lower);
		r.angB = mk4(angCompB.x, angCompB.y, angCompB.z, upper);
		r.appliedImpulse = 0.f;
		r.bodyA = c.rbA;
		r.bodyB = c.rbB;
		r.pad = 0;
		rows[j] = r;
	}
}

// resolveSingleConstraintRowGeneric (b3GpuPgsConstraintSolver.cpp:575-604, jointSolver.cl:262-300); deltas through L2
B3_D void resolveRow(const JointArgs& a, JointRow* row)
{
	JointRow r = *row;
	const float invMassA = a.pose[2 * r.bodyA].w, invMassB = a.pose[2 * r.bodyB].w;
	float4 dLinA = __ldcg(&a.delta[2 * r.bodyA]), dAngA = __ldcg(&a.delta[2 * r.bodyA + 1]);
	float4 dLinB = __ldcg(&a.delta[2 * r.bodyB]), dAngB = __ldcg(&a.delta[2 * r.bodyB + 1]);
	const float jacDiagABInv = r.normal.w, rhs = r.rel1.w, cfm = r.rel2.w, lower = r.angA.w, upper = r.angB.w;
	float deltaImpulse = rhs - r.appliedImpulse * cfm;
	const float deltaVel1Dotn = dot3(r.normal, dLinA) + dot3(r.rel1, dAngA);
	const float deltaVel2Dotn = -dot3(r.normal, dLinB) + dot3(r.rel2, dAngB);
	deltaImpulse -= deltaVel1Dotn * jacDiagABInv;
	deltaImpulse -= deltaVel2Dotn * jacDiagABInv;
	const float sum = r.appliedImpulse + deltaImpulse;
	if (sum < lower)
	{
		deltaImpulse = lower - r.appliedImpulse;
		r.appliedImpulse = lower;
	}
	else if (sum > upper)
	{
		deltaImpulse = upper - r.appliedImpulse;
		r.appliedImpulse = upper;
	}
	else
		r.appliedImpulse = sum;
	row->appliedImpulse = r.appliedImpulse;
	// internalApplyImpulse (linearFactor = angularFactor = 1): dv += (normal * invMass) * impulse, dw += angularComponent * impulse
	dLinA = add3(dLinA, scale3(scale3(r.normal, invMassA), deltaImpulse));
	dAngA = add3(dAngA, scale3(r.angA, deltaImpulse));
	dLinB = add3(dLinB, scale3(scale3(neg3(r.normal), invMassB), deltaImpulse));
	dAngB = add3(dAngB, scale3(r.angB, deltaImpulse));
	__stcg(&a.delta[2 * r.bodyA], dLinA);
	__stcg(&a.delta[2 * r.bodyA + 1], dAngA);
	__stcg(&a.delta[2 * r.bodyB], dLinB);
	__stcg(&a.delta[2 * r.bodyB + 1], dAngB);
}

B3_D void solveOne(const JointArgs& a, int slot)
{
	const int ci = a.order[slot];
	if (!(a.joints[ci].flags & B3B200_CONSTRAINT_FLAG_ENABLED)) return;
	const int n = a.numRows[ci];
	JointRow* rows = a.rows + (size_t)ci * J_MAX_ROWS;
	for (int j = 0; j < n; j++) resolveRow(a, &rows[j]);
}

__global__ void __launch_bounds__(J_THREADS) jointSetupKernel(JointArgs a)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a.numJoints) setupOne(a, i);
}

// all iterations x batches in one CTA (every batch has at most blockDim.x constraints)
__global__ void __launch_bounds__(1024) jointSolveSingleCtaKernel(JointArgs a)
{
	for (int it = 0; it < a.iterations; it++)
		for (int b = 0; b < a.numBatches; b++)
		{
			const int begin = a.batchOffset[b], end = a.batchOffset[b + 1];
			for (int s = begin + (int)threadIdx.x; s < end; s += (int)blockDim.x) solveOne(a, s);
			__syncthreads();
		}
}

__global__ void __launch_bounds__(J_THREADS) jointSolveBatchKernel(JointArgs a, int begin, int end)
{
	const int s = begin + blockIdx.x * blockDim.x + threadIdx.x;
	if (s < end) solveOne(a, s);
}

// breakViolatedConstraintsKernel (jointSolver.cl:352-372)
__global__ void __launch_bounds__(J_THREADS) jointBreakKernel(JointArgs a)
{
	const int ci = blockIdx.x * blockDim.x + threadIdx.x;
	if (ci >= a.numJoints) return;
	const int n = a.numRows[ci];
	const float thr = a.joints[ci].breakingImpulseThreshold;
	for (int j = 0; j < n; j++)
		if (fabsf(a.rows[(size_t)ci * J_MAX_ROWS + j].appliedImpulse) >= thr) a.joints[ci].flags = 0;
}

// writeBackVelocitiesKernel (jointSolver.cl:631-649): v += dv for dynamic bodies
__global__ void __launch_bounds__(J_THREADS) jointWriteBackKernel(JointArgs a)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.numBodies) return;
	if (a.pose[2 * i].w == 0.f) return;
	const float4 dl = a.delta[2 * i], da = a.delta[2 * i + 1];
	if (dl.x == 0.f && dl.y == 0.f && dl.z == 0.f && da.x == 0.f && da.y == 0.f && da.z == 0.f) return;  // untouched by any joint
	float4 l = a.vel[2 * i], w = a.vel[2 * i + 1];
	a.vel[2 * i] = mk4(l.x + dl.x, l.y + dl.y, l.z + dl.z, l.w);
	a.vel[2 * i + 1] = mk4(w.x + da.x, w.y + da.y, w.z + da.z, w.w);
}
}  // namespace

// sortConstraintByBatch3 (b3GpuPgsConstraintSolver.cpp:778-905) with simdWidth = numConstraints + 1: repeated passes over
// the not yet batched constraints in order; a constraint is taken when none of its dynamic bodies is in the pass yet.
static void batchJoints(World* w)
{
	const int n = (int)w->joints.size();
	w->jointOrder.resize(n);
	w->jointBatchOffset.clear();
	w->jointBatchOffset.push_back(0);
	std::vector<int> cs(n);
	for (int i = 0; i < n; i++) cs[i] = i;
	std::vector<unsigned char> used(w->bodies.size(), 0);
	std::vector<int> touched;
	int numValid = 0;
	while (numValid < n)
	{
		for (int b : touched) used[b] = 0;
		touched.clear();
		for (int i = numValid; i < n; i++)
		{
			const b3b200_generic_constraint& c = w->joints[cs[i]];
			const bool aStatic = w->bodies[c.rbA].invMass == 0.f, bStatic = w->bodies[c.rbB].invMass == 0.f;
			bool unavailable = !aStatic && used[c.rbA];
			if (!unavailable) unavailable = !bStatic && used[c.rbB];
			if (unavailable) continue;
			if (!aStatic)
			{
				used[c.rbA] = 1;
				touched.push_back(c.rbA);
			}
			if (!bStatic)
			{
				used[c.rbB] = 1;
				touched.push_back(c.rbB);
			}
			std::swap(cs[i], cs[numValid]);
			numValid++;
		}
		w->jointBatchOffset.push_back(numValid);
	}
	for (int i = 0; i < n; i++) w->jointOrder[i] = cs[i];
	w->jointBatchesDirty = false;
}

int launchSolveJoints(World* w)
{
	const int n = (int)w->joints.size();
	if (n == 0) return 0;
	cudaStream_t s = w->stream;
	if (w->jointsDirty)
	{
		// constraints created / removed on the host since the last step (the device copy carries the broken flags)
		B3_TRY(w->dJoints.reserve((size_t)n));
		B3_CUDA_CHECK(cudaMemcpyAsync(w->dJoints.ptr, w->joints.data(), sizeof(b3b200_generic_constraint) * (size_t)n, cudaMemcpyHostToDevice, s));
		w->jointsDirty = false;
	}
	if (w->jointBatchesDirty)
	{
		batchJoints(w);
		B3_TRY(w->dJointOrder.reserve((size_t)n));
		B3_TRY(w->dJointBatchOffset.reserve(w->jointBatchOffset.size()));
		B3_CUDA_CHECK(cudaMemcpyAsync(w->dJointOrder.ptr, w->jointOrder.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s));
		B3_CUDA_CHECK(cudaMemcpyAsync(w->dJointBatchOffset.ptr, w->jointBatchOffset.data(), sizeof(int) * w->jointBatchOffset.size(), cudaMemcpyHostToDevice, s));
		B3_CUDA_CHECK(cudaStreamSynchronize(s));  // the host vectors may be rebuilt before the copies ran otherwise
	}
	B3_TRY(w->dJointRows.reserve((size_t)n * J_MAX_ROWS * (sizeof(JointRow) / sizeof(float4))));
	B3_TRY(w->dJointNumRows.reserve((size_t)n));
	B3_TRY(w->dJointDelta.reserve(2 * (size_t)std::max(w->numBodies, 1)));
	B3_CUDA_CHECK(cudaMemsetAsync(w->dJointDelta.ptr, 0, sizeof(float4) * 2 * (size_t)w->numBodies, s));
	JointArgs a;
	a.joints = w->dJoints.ptr;
	a.numJoints = n;
	a.order = w->dJointOrder.ptr;
	a.batchOffset = w->dJointBatchOffset.ptr;
	a.numBatches = (int)w->jointBatchOffset.size() - 1;
	a.rows = reinterpret_cast<JointRow*>(w->dJointRows.ptr);
	a.numRows = w->dJointNumRows.ptr;
	a.delta = w->dJointDelta.ptr;
	a.pose = w->dPose.ptr;
	a.vel = w->dVel.ptr;
	a.inertias = w->dInertias.ptr;
	a.numBodies = w->numBodies;
	a.iterations = 4;       // b3GpuPgsConstraintSolver::solveJoints :932
	a.fps = 60.f;           // 1 / m_timeStep, m_timeStep = 1/60 (:931)
	a.erp = 0.2f;           // b3ContactSolverInfo defaults (b3ContactSolverInfo.h:66-74)
	a.cfm = 0.f;
	a.damping = 1.0f;
	jointSetupKernel<<<divUp(n, J_THREADS), J_THREADS, 0, s>>>(a);
	B3_LAUNCH_CHECK();
	int maxBatch = 0;
	for (size_t b = 0; b + 1 < w->jointBatchOffset.size(); b++) maxBatch = std::max(maxBatch, w->jointBatchOffset[b + 1] - w->jointBatchOffset[b]);
	if (maxBatch <= 4096)
	{
		jointSolveSingleCtaKernel<<<1, maxBatch <= 256 ? 256 : 1024, 0, s>>>(a);
		B3_LAUNCH_CHECK();
	}
	else
	{
		for (int it = 0; it < a.iterations; it++)
			for (int b = 0; b < a.numBatches; b++)
			{
				const int begin = w->jointBatchOffset[b], end = w->jointBatchOffset[b + 1];
				jointSolveBatchKernel<<<divUp(end - begin, J_THREADS), J_THREADS, 0, s>>>(a, begin, end);
				B3_LAUNCH_CHECK();
			}
	}
	jointBreakKernel<<<divUp(n, J_THREADS), J_THREADS, 0, s>>>(a);
	B3_LAUNCH_CHECK();
	jointWriteBackKernel<<<divUp(w->numBodies, J_THREADS), J_THREADS, 0, s>>>(a);
	B3_LAUNCH_CHECK();
	w->soaDirty = true;
	return 0;
}

}  // namespace b3b200

using namespace b3b200;

static int createJoint(b3b200_world* w, int type, int bodyA, int bodyB, const float* pA, const float* pB, const float* rel, float thr)
{
	if (!w || !pA || !pB || bodyA < 0 || bodyB < 0 || bodyA >= (int)w->bodies.size() || bodyB >= (int)w->bodies.size())
	{
		setLastError("createConstraint: invalid argument");
		return -1;
	}
	b3b200_generic_constraint c;
	memset(&c, 0, sizeof(c));
	c.constraintType = type;
	c.rbA = bodyA;
	c.rbB = bodyB;
	c.breakingImpulseThreshold = thr;
	c.pivotInA.x = pA[0];
	c.pivotInA.y = pA[1];
	c.pivotInA.z = pA[2];
	c.pivotInB.x = pB[0];
	c.pivotInB.y = pB[1];
	c.pivotInB.z = pB[2];
	if (rel)
	{
		c.relTargetAB.x = rel[0];
		c.relTargetAB.y = rel[1];
		c.relTargetAB.z = rel[2];
		c.relTargetAB.w = rel[3];
	}
	c.flags = B3B200_CONSTRAINT_FLAG_ENABLED;
	c.uid = w->jointUid++;
	if (b3b200_get_joints(w, nullptr, 0, nullptr) < 0) return -1;  // refresh the host copy (broken flags) before it is re-sent
	w->joints.push_back(c);
	w->jointsDirty = true;
	w->jointBatchesDirty = true;
	return c.uid;
}

extern "C" int b3b200_create_p2p_constraint(b3b200_world* w, int bodyA, int bodyB, const float* pivotInA3, const float* pivotInB3, float breakingThreshold)
{
	return createJoint(w, B3B200_CONSTRAINT_P2P, bodyA, bodyB, pivotInA3, pivotInB3, nullptr, breakingThreshold);
}

extern "C" int b3b200_create_fixed_constraint(b3b200_world* w, int bodyA, int bodyB, const float* pivotInA3, const float* pivotInB3, const float* relTargetAB4,
											  float breakingThreshold)
{
	if (!relTargetAB4)
	{
		setLastError("createFixedConstraint: invalid argument");
		return -1;
	}
	return createJoint(w, B3B200_CONSTRAINT_FIXED, bodyA, bodyB, pivotInA3, pivotInB3, relTargetAB4, breakingThreshold);
}

extern "C" int b3b200_remove_constraint(b3b200_world* w, int uid)
{
	if (!w) return B3B200_ERR_INVALID;
	if (b3b200_get_joints(w, nullptr, 0, nullptr) < 0) return B3B200_ERR_CUDA;
	// removeConstraintByUid (b3GpuRigidBodyPipeline.cpp:158-185): swap with the last, pop
	for (size_t i = 0; i < w->joints.size(); i++)
		if (w->joints[i].uid == uid)
		{
			std::swap(w->joints[i], w->joints.back());
			w->joints.pop_back();
			w->jointsDirty = true;
			w->jointBatchesDirty = true;
			return 0;
		}
	return 0;  // unknown uid: the reference ignores it too
}

extern "C" int b3b200_num_constraints(b3b200_world* w) { return w ? (int)w->joints.size() : B3B200_ERR_INVALID; }

// copyConstraintsToHost (b3GpuRigidBodyPipeline.cpp:593-596): the device copy carries the flags of broken joints
extern "C" int b3b200_get_joints(b3b200_world* w, b3b200_generic_constraint* dst, int capacity, int* count)
{
	if (!w) return B3B200_ERR_INVALID;
	const int n = (int)w->joints.size();
	if (n && !w->jointsDirty && w->device >= 0 && w->dJoints.ptr)
	{
		B3_CUDA_CHECK(cudaSetDevice(w->device));
		B3_CUDA_CHECK(cudaMemcpyAsync(w->joints.data(), w->dJoints.ptr, sizeof(b3b200_generic_constraint) * (size_t)n, cudaMemcpyDeviceToHost, w->stream));
		B3_CUDA_CHECK(cudaStreamSynchronize(w->stream));
	}
	if (count) *count = n;
	if (dst)
		for (int i = 0; i < n && i < capacity; i++) dst[i] = w->joints[i];
	return 0;
}

extern "C" int b3b200_solve_joints(b3b200_world* w)
{
	if (w) w->dropStepGraphs();  // may change what a step launches
	if (!w || w->device < 0 || !w->uploaded) return B3B200_ERR_STATE;
	B3_CUDA_CHECK(cudaSetDevice(w->device));
	return launchSolveJoints(w);
}
