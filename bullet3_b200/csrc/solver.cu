// solver.cu -- batched PGS contact solver.
//
// Replaces b3GpuPgsContactSolver::solveContacts (b3GpuPgsContactSolver.cpp:568-1103):
// 6 radix sorts + a serial per-cell batching thread (batchingKernelsNew.cl:144-231)
// + contact->constraint + 2*I*8 launches of 32 work-groups, with host
// synchronisation after every phase.  Here:
//
//  setup    ONE persistent cooperative kernel:
//           (1) graph colouring of the contact graph by priority rounds
//               (Jones-Plassmann): every contact has a 64-bit priority that is a
//               hash of (bodyA, bodyB, childA, childB); per round each uncoloured
//               contact posts its priority on its dynamic bodies with atomicMax,
//               and the contact that is top on both of its bodies takes the
//               lowest colour not yet used on either body (two 64-bit colour
//               masks per body = 128 colours = B3_MAX_NUM_BATCHES, b3Solver.h:39).
//               The result equals a sequential first-fit colouring in
//               descending priority order, i.e. it is deterministic whatever
//               order the narrowphase appended the contacts in, and the CPU
//               oracle reproduces it exactly ("same batching").
//           (2) colour histogram -> batch offsets,
//           (3) contact -> constraint rows (setConstraint4,
//               b3ConvertConstraint4.h:62-148) written straight into batch order.
//  iterate  ONE persistent cooperative kernel: for every iteration, every batch
//           in order (grid barrier in between), all normal rows; then the same
//           for friction -- the order of the reference's global-batch mode
//           (gUseLargeBatches, solveContactConstraintBatchSizes,
//           b3GpuPgsContactSolver.cpp:262-311) and of its host twin
//           (solveContact<false>/solveFriction, b3Solver.cpp:187-329).
//           Velocities live in the 32-byte-per-body SoA array (8 MB at 256k
//           bodies: L2-resident across all 2*I*batches phases).
#include "internal.h"

namespace b3b200
{
constexpr int SOLVER_THREADS = 512;
constexpr int ITER_THREADS = 512;
constexpr int TAIL_ROWS = ITER_THREADS;  // batches up to one row per thread are cheaper to solve in one CTA (measured 1.7 us
                                           // per phase) than to pay a grid barrier for (3.5 us); two rows per thread cost 5.5 us
constexpr int MAX_ROUNDS = 1024;
constexpr int TAIL_COLOUR = 2048;  // colouring rounds with at most this many contacts left run in CTA 0 alone

// ---------------------------------------------------------------- grid barrier
// One monotonically increasing arrival counter (zeroed by the host before the launch).
// Thread 0 of every CTA arrives with a gpu-scope RELEASE add and spins with gpu-scope
// ACQUIRE loads until all CTAs of this generation have arrived; bar.sync on either side
// extends the ordering to the rest of the CTA (PTX memory model: causality order through
// the CTA barrier, release/acquire are cumulative).  All CTAs are co-resident
// (cooperative launch).
// MEASURED CAVEAT: the acquire by thread 0 does not stop the OTHER threads' plain loads from hitting
// stale lines in the SM's L1 (body colour masks read with ld.global gave invalid colourings in ~30 %
// of the runs once the rounds became short).  Every value another CTA may have written during this
// kernel is therefore read with ld.global.cg (__ldcg), volatile or an atomic -- never a plain load.
struct GridBarrier
{
	unsigned int* counter;
	unsigned int numBlocks;
	unsigned int target;
	B3_D void init(unsigned int* c, unsigned int nb)
	{
		counter = c;
		numBlocks = nb;
		target = 0;
	}
	// split form: arrive() publishes this CTA's writes and signals, wait() blocks until every CTA has arrived;
	// independent loads may be issued in between
	unsigned int seen;
	B3_D void arrive()
	{
		target += numBlocks;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			// fire and forget: nobody needs the old value, and waiting for it would put one more L2 round trip in front of
			// the first poll
			asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
			seen = target - 1u;
		}
	}
	B3_D void wait()
	{
		if (threadIdx.x == 0)
		{
			// poll with relaxed loads and acquire ONCE at the end: an acquire load compiles to LD + CCTL.IVALL, i.e. every
			// poll would throw away the SM's whole L1
			unsigned int v = seen;
			while ((int)(v - target) < 0)
			{
				asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
			}
			asm volatile("fence.acq_rel.gpu;" ::: "memory");
		}
		__syncthreads();
	}
	B3_D void sync()
	{
		target += numBlocks;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
			unsigned int v = target - 1u;
			while ((int)(v - target) < 0)
			{
				asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
			}
			asm volatile("fence.acq_rel.gpu;" ::: "memory");
		}
		__syncthreads();
	}
};

B3_HD unsigned int hashContact(int a, int b, int ca, int cb)
{
	unsigned int h = (unsigned int)a * 0x9E3779B1u;
	h ^= (unsigned int)b * 0x85EBCA77u + 0x165667B1u + (h << 6) + (h >> 2);
	h ^= (unsigned int)ca * 0xC2B2AE3Du + (h << 6) + (h >> 2);
	h ^= (unsigned int)cb * 0x27D4EB2Fu + (h << 6) + (h >> 2);
	h ^= h >> 16;
	h *= 0x85EBCA6Bu;
	h ^= h >> 13;
	h *= 0xC2B2AE35u;
	h ^= h >> 16;
	return h;
}

struct SetupArgs
{
	b3b200_contact4* contacts;
	unsigned int* ctr;
	const float4* pose;
	const float4* vel;
	const b3b200_inertia* inertias;
	b3b200_constraint4* constraints;
	unsigned long long* bodyMask;  // 2 per body
	unsigned long long* bodyPrio;  // 1 per body
	int* contactColour;
	unsigned int* batchCount;   // MAX_BATCHES
	unsigned int* batchOffset;  // MAX_BATCHES + 1
	unsigned int* batchCursor;  // MAX_BATCHES
	unsigned int* remaining;    // MAX_ROUNDS
	unsigned int* colourList;   // 2 x colourListStride: uncoloured contacts of the current / next round
	int colourListStride;
	unsigned int* bar;
	int numBodies;
	int staticIdx;
	int colouring;  // 0 = Jones-Plassmann rounds (reproducible for a given contact array), 1 = single-pass first fit with atomics
	float dt, positionDrift, positionConstraintCoeff;
};

B3_D float4 matRowMul(const float4& r0, const float4& r1, const float4& r2, const float4& v)
{
	return mk4(dot3(r0, v), dot3(r1, v), dot3(r2, v));
}

// calcJacCoeff (b3ConvertConstraint4.h:50-60)
B3_D float calcJacCoeff(const float4& angular0, const float4& angular1, float invMass0, const float4* I0, float invMass1, const float4* I1)
{
	float jmj0 = invMass0;
	float jmj1 = dot3(matRowMul(I0[0], I0[1], I0[2], angular0), angular0);
	float jmj2 = invMass1;
	float jmj3 = dot3(matRowMul(I1[0], I1[1], I1[2], angular1), angular1);
	return -1.f / (jmj0 + jmj1 + jmj2 + jmj3);
}
B3_D float calcRelVel(const float4& l0, const float4& l1, const float4& a0, const float4& a1, const float4& linVel0, const float4& angVel0,
					  const float4& linVel1, const float4& angVel1)
{
	return dot3(l0, linVel0) + dot3(a0, angVel0) + dot3(l1, linVel1) + dot3(a1, angVel1);
}
// b3PlaneSpace1 (b3ConvertConstraint4.h:5-34)
B3_D void planeSpace1(const float4& n, float4& p, float4& q)
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

// setConstraint4 (b3ConvertConstraint4.h:62-148)
B3_D void buildConstraint(const SetupArgs& s, const b3b200_contact4* __restrict__ src, int colour, b3b200_constraint4* __restrict__ dst)
{
	const float4* cw = reinterpret_cast<const float4*>(src);
	float4 wp[4] = {cw[0], cw[1], cw[2], cw[3]};
	const float4 nrm = cw[4];
	const int4 ids = reinterpret_cast<const int4*>(src)[5];
	const int aIdx = abs(ids.z), bIdx = abs(ids.w);
	const float4 posA = s.pose[2 * aIdx], posB = s.pose[2 * bIdx];
	const float4 linVelA = s.vel[2 * aIdx], angVelA = s.vel[2 * aIdx + 1];
	const float4 linVelB = s.vel[2 * bIdx], angVelB = s.vel[2 * bIdx + 1];
	const float invMassA = posA.w, invMassB = posB.w;
	// quirk kept from the reference: rows are built with the LOCAL initial inverse inertia
	// (solverSetup.cl:254,260 / b3Solver.cpp:911,917), solved with the world one.
	const float4* IA = reinterpret_cast<const float4*>(&s.inertias[aIdx].initInvInertia);
	const float4* IB = reinterpret_cast<const float4*>(&s.inertias[bIdx].initInvInertia);
	float4 ia[3] = {__ldg(IA), __ldg(IA + 1), __ldg(IA + 2)};
	float4 ib[3] = {__ldg(IB), __ldg(IB + 1), __ldg(IB + 2)};

	const float dtInv = 1.f / s.dt;
	const float npoints = nrm.w;
	float jac[4], bb[4];
	const float4 n = mk4(nrm.x, nrm.y, nrm.z);
	const float4 nn = neg3(n);
#pragma unroll
	for (int ic = 0; ic < 4; ic++)
	{
		float4 r0 = sub3(wp[ic], posA);
		float4 r1 = sub3(wp[ic], posB);
		if ((float)ic >= npoints)
		{
			jac[ic] = 0.f;
			bb[ic] = 0.f;
			continue;
		}
		float4 angular0 = cross3(r0, n);
		float4 angular1 = neg3(cross3(r1, n));
		jac[ic] = calcJacCoeff(angular0, angular1, invMassA, ia, invMassB, ib);
		float relVelN = calcRelVel(n, nn, angular0, angular1, linVelA, angVelA, linVelB, angVelB);
		float e = 0.f;
		float b = e * relVelN;
		b += (wp[ic].w + s.positionDrift) * s.positionConstraintCoeff * dtInv;
		bb[ic] = b;
	}
	float fjac[2] = {0.f, 0.f};
	float4 center = mk4(0, 0, 0);
	if (npoints > 0)
	{
		for (int i = 0; (float)i < npoints && i < 4; i++)
		{
			center.x += wp[i].x;
			center.y += wp[i].y;
			center.z += wp[i].z;
		}
		float inv = 1.0f / (float)npoints;
		center.x *= inv;
		center.y *= inv;
		center.z *= inv;
		float4 t0, t1;
		planeSpace1(n, t0, t1);
		float4 r0 = sub3(center, posA), r1 = sub3(center, posB);
		{
			float4 a0 = cross3(r0, t0), a1 = neg3(cross3(r1, t0));
			fjac[0] = calcJacCoeff(a0, a1, invMassA, ia, invMassB, ib);
		}
		{
			float4 a0 = cross3(r0, t1), a1 = neg3(cross3(r1, t1));
			fjac[1] = calcJacCoeff(a0, a1, invMassA, ia, invMassB, ib);
		}
	}
	float4* dw = reinterpret_cast<float4*>(dst);
	dw[0] = mk4(nrm.x, nrm.y, nrm.z, 0.7f);
#pragma unroll
	for (int i = 0; i < 4; i++) dw[1 + i] = ((float)i < npoints) ? wp[i] : mk4(0, 0, 0, 0);
	dw[5] = center;
	dw[6] = mk4(jac[0], jac[1], jac[2], jac[3]);
	dw[7] = mk4(bb[0], bb[1], bb[2], bb[3]);
	dw[8] = mk4(0, 0, 0, 0);              // appliedRambdaDt
	dw[9] = mk4(fjac[0], fjac[1], 0, 0);  // fJacCoeffInv, fAppliedRambdaDt
	int4 tail;
	tail.x = aIdx;
	tail.y = bIdx;
	tail.z = colour;
	tail.w = 0;
	reinterpret_cast<int4*>(dst)[10] = tail;
}

// One colouring round, step 1: an uncoloured contact posts its priority on its dynamic bodies.
B3_D void colourClaim(const SetupArgs& s, int c)
{
	const int4 ids = reinterpret_cast<const int4*>(&s.contacts[c])[5];
	const int4 ch = reinterpret_cast<const int4*>(&s.contacts[c])[6];
	const int a = abs(ids.z), b = abs(ids.w);
	const bool aStatic = ids.z < 0 || ids.z == s.staticIdx;
	const bool bStatic = ids.w < 0 || ids.w == s.staticIdx;
	const unsigned long long prio = ((unsigned long long)hashContact(a, b, ch.x, ch.y) << 32) | (unsigned long long)(c + 1);
	if (!aStatic) atomicMax(&s.bodyPrio[a], prio);
	if (!bStatic) atomicMax(&s.bodyPrio[b], prio);
}

// Step 2: the contact that is top on both of its bodies takes the lowest colour free on both.  Returns the colour, -1 when
// the contact stays uncoloured for the next round, -2 when it cannot be coloured at all (the caller counts the colours,
// warp-aggregated).
B3_D int colourTry(const SetupArgs& s, int c)
{
	const int4 ids = reinterpret_cast<const int4*>(&s.contacts[c])[5];
	const int4 ch = reinterpret_cast<const int4*>(&s.contacts[c])[6];
	const int a = abs(ids.z), b = abs(ids.w);
	const bool aStatic = ids.z < 0 || ids.z == s.staticIdx;
	const bool bStatic = ids.w < 0 || ids.w == s.staticIdx;
	const unsigned long long prio = ((unsigned long long)hashContact(a, b, ch.x, ch.y) << 32) | (unsigned long long)(c + 1);
	volatile unsigned long long* vp = s.bodyPrio;
	const bool top = (aStatic || vp[a] == prio) && (bStatic || vp[b] == prio);
	if (!top) return -1;
	// (masks written by other CTAs in earlier rounds: read through L2)
	unsigned long long m0 = 0ull, m1 = 0ull;
	if (!aStatic)
	{
		m0 |= __ldcg(&s.bodyMask[2 * a]);
		m1 |= __ldcg(&s.bodyMask[2 * a + 1]);
	}
	if (!bStatic)
	{
		m0 |= __ldcg(&s.bodyMask[2 * b]);
		m1 |= __ldcg(&s.bodyMask[2 * b + 1]);
	}
	int colour = -2;
	if (~m0)
		colour = __ffsll((long long)~m0) - 1;
	else if (~m1)
		colour = 64 + __ffsll((long long)~m1) - 1;
	if (colour < 0)
	{
		// more than B3_MAX_NUM_BATCHES colours at one body: the reference errors out
		// ("batchIdx>=B3_MAX_NUM_BATCHES", b3GpuPgsContactSolver.cpp:1497-1502); here the
		// contact is left out of this step's solve and the overflow flag is raised.
		s.contactColour[c] = -2;
		s.contacts[c].batchIdx = -2;
		if (!aStatic) s.bodyPrio[a] = 0ull;
		if (!bStatic) s.bodyPrio[b] = 0ull;
		atomicOr(&s.ctr[CTR_OVERFLOW], (unsigned int)OVF_BATCHES);
		return -2;
	}
	const unsigned long long bit = 1ull << (colour & 63);
	const int word = colour >> 6;
	if (!aStatic)
	{
		__stcg(&s.bodyMask[2 * a + word], __ldcg(&s.bodyMask[2 * a + word]) | bit);
		s.bodyPrio[a] = 0ull;
	}
	if (!bStatic)
	{
		__stcg(&s.bodyMask[2 * b + word], __ldcg(&s.bodyMask[2 * b + word]) | bit);
		s.bodyPrio[b] = 0ull;
	}
	s.contactColour[c] = colour;
	s.contacts[c].batchIdx = colour;
	return colour;
}

// Single-pass colouring: every contact takes the lowest colour that is free on both of its dynamic bodies by setting the
// colour's bit in the bodies' masks with atomicOr -- whoever flips a bit from 0 to 1 owns that (body, colour) -- and
// retries with fresh masks when it loses a race.  The bodies are taken in index order, so two contacts can never hold
// one bit each and wait for the other's (no livelock); a contact that loses on its second body gives the first bit back.
// No rounds and no grid barrier: a body's contacts resolve their races in a few L2 round trips.  The colours depend on
// the order in which the races resolve, i.e. they are not reproducible from run to run (the Jones-Plassmann path is).
B3_D int colourFirstFit(const SetupArgs& s, int c)
{
	const int4 ids = reinterpret_cast<const int4*>(&s.contacts[c])[5];
	const int a = abs(ids.z), b = abs(ids.w);
	const bool aStatic = ids.z < 0 || ids.z == s.staticIdx;
	const bool bStatic = ids.w < 0 || ids.w == s.staticIdx;
	int body[2];
	int nb = 0;
	if (!aStatic) body[nb++] = a;
	if (!bStatic) body[nb++] = b;
	if (nb == 2 && body[0] > body[1])
	{
		const int t = body[0];
		body[0] = body[1];
		body[1] = t;
	}
	int colour = 0;
	for (;;)
	{
		unsigned long long m0 = 0ull, m1 = 0ull;
		for (int k = 0; k < nb; k++)
		{
			m0 |= __ldcg(&s.bodyMask[2 * body[k]]);
			m1 |= __ldcg(&s.bodyMask[2 * body[k] + 1]);
		}
		colour = -2;
		if (~m0)
			colour = __ffsll((long long)~m0) - 1;
		else if (~m1)
			colour = 64 + __ffsll((long long)~m1) - 1;
		if (colour < 0)
		{
			// more than B3_MAX_NUM_BATCHES colours at one body (see colourTry)
			s.contactColour[c] = -2;
			s.contacts[c].batchIdx = -2;
		s.contacts[c].batchIdx = -2;
			atomicOr(&s.ctr[CTR_OVERFLOW], (unsigned int)OVF_BATCHES);
			return -2;
		}
		const unsigned long long bit = 1ull << (colour & 63);
		const int word = colour >> 6;
		int got = 0;
		for (; got < nb; got++)
			if (atomicOr(&s.bodyMask[2 * body[got] + word], bit) & bit) break;
		if (got == nb) break;
		for (int k = 0; k < got; k++) atomicAnd(&s.bodyMask[2 * body[k] + word], ~bit);
	}
	s.contactColour[c] = colour;
	s.contacts[c].batchIdx = colour;
	return colour;  // the caller counts it (warp-aggregated)
}

// counters[key] += 1 for every lane with key >= 0, one atomic per distinct key in the warp (a few hundred thousand contacts
// share ~20 colours: one atomic each would serialise on ~20 addresses).  Returns the lane's own slot.  Whole warp calls.
B3_D unsigned int warpCountByKey(unsigned int* counters, int key, int lane)
{
	const unsigned int peers = __match_any_sync(0xffffffffu, key);
	unsigned int base = 0;
	const int leader = __ffs(peers) - 1;
	if (key >= 0 && lane == leader) base = atomicAdd(&counters[key], (unsigned int)__popc(peers));
	base = __shfl_sync(0xffffffffu, base, leader);
	return base + (unsigned int)__popc(peers & ((1u << lane) - 1u));
}

#ifdef B3_SETUP_TIMING
#define B3_PROBE(tag) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); printf("setup %s %llu round %d\n", tag, t, round); } } while (0)
#else
#define B3_PROBE(tag)
#endif
__global__ void __launch_bounds__(SOLVER_THREADS) solverSetupKernel(SetupArgs s)
{
	GridBarrier bar;
	bar.init(s.bar, gridDim.x);
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int stride = gridDim.x * blockDim.x;
	const int nContacts = (int)s.ctr[CTR_CONTACTS];
	int round = 0;
	B3_PROBE("start");

	// ---- phase 0: clear
	for (int i = tid; i < s.numBodies; i += stride)
	{
		s.bodyMask[2 * i] = 0ull;
		s.bodyMask[2 * i + 1] = 0ull;
		s.bodyPrio[i] = 0ull;
	}
	for (int i = tid; i < nContacts; i += stride) s.contactColour[i] = -1;
	for (int i = tid; i < MAX_BATCHES; i += stride)
	{
		s.batchCount[i] = 0;
		s.batchCursor[i] = 0;
	}
	for (int i = tid; i < MAX_ROUNDS; i += stride) s.remaining[i] = 0;
	bar.sync();

	B3_PROBE("cleared");
	// ---- phase 1: colouring rounds
	// Round 0 walks all contacts; every later round walks the compacted list of the contacts the previous round left
	// uncoloured (ping-pong halves of colourList; remaining[round] is both the list length and the stop criterion).
	// Once the list is short (<= TAIL_COLOUR) CTA 0 finishes the remaining rounds alone between __syncthreads():
	// two grid barriers per round cost more than such a round.
	int count = nContacts;
	const unsigned int* cur = nullptr;
	const int lane = threadIdx.x & 31;
	__shared__ unsigned int sNext;
	if (s.colouring == 1)
	{
		for (int base = tid - lane; base < nContacts; base += stride)
		{
			const int c = base + lane;
			const int colour = c < nContacts ? colourFirstFit(s, c) : -1;
			__syncwarp();
			warpCountByKey(s.batchCount, colour, lane);
		}
		bar.sync();
		count = 0;
	}
	for (; count > 0 && round < MAX_ROUNDS; round++)
	{
		for (int i = tid; i < count; i += stride) colourClaim(s, cur ? (int)__ldcg(&cur[i]) : i);
		bar.sync();
		unsigned int* nxt = s.colourList + (size_t)(round & 1) * (size_t)s.colourListStride;
		for (int base = tid - lane; base < count; base += stride)
		{
			const int i = base + lane;
			const int c = i < count ? (cur ? (int)__ldcg(&cur[i]) : i) : 0;
			const int got = i < count ? colourTry(s, c) : -2;
			const bool left = got == -1;
			__syncwarp();
			warpCountByKey(s.batchCount, got, lane);
			// the still uncoloured contacts of this warp go to the next round's list
			const unsigned int m = __ballot_sync(0xffffffffu, left);
			if (m)
			{
				unsigned int slot = 0;
				if (lane == 0) slot = atomicAdd(&s.remaining[round], (unsigned int)__popc(m));
				slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(m & ((1u << lane) - 1u));
				if (left) nxt[slot] = (unsigned int)c;
			}
		}
		bar.sync();
		volatile unsigned int* vr = s.remaining;
		count = (int)vr[round];
		cur = nxt;
		if (count == 0) break;
		if (count <= TAIL_COLOUR)
		{
			if (blockIdx.x == 0)
			{
				while (count > 0 && round + 1 < MAX_ROUNDS)
				{
					round++;
					for (int i = threadIdx.x; i < count; i += blockDim.x) colourClaim(s, (int)__ldcg(&cur[i]));
					if (threadIdx.x == 0) sNext = 0;
					__syncthreads();
					nxt = s.colourList + (size_t)(round & 1) * (size_t)s.colourListStride;
					for (int base = (int)threadIdx.x - lane; base < count; base += blockDim.x)
					{
						const int i = base + lane;
						const int c = i < count ? (int)__ldcg(&cur[i]) : 0;
						const int got = i < count ? colourTry(s, c) : -2;
						const bool left = got == -1;
						__syncwarp();
						warpCountByKey(s.batchCount, got, lane);
						const unsigned int m = __ballot_sync(0xffffffffu, left);
						if (m)
						{
							unsigned int slot = 0;
							if (lane == 0) slot = atomicAdd(&sNext, (unsigned int)__popc(m));
							slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(m & ((1u << lane) - 1u));
							if (left) nxt[slot] = (unsigned int)c;
						}
					}
					__syncthreads();
					count = (int)sNext;
					cur = nxt;
					__syncthreads();
				}
			}
			bar.sync();
			break;
		}
	}

	B3_PROBE("coloured");
	// ---- phase 2: batch offsets (one warp)
	if (blockIdx.x == 0 && threadIdx.x < 32)
	{
		unsigned int run = 0;
		int numBatches = 0;
		for (int base = 0; base < MAX_BATCHES; base += 32)
		{
			// every batch is padded to a multiple of 32 slots, so that a warp-row of the iteration
			// kernels never straddles two batches (padding slots are marked invalid below)
			const unsigned int raw = __ldcg(&s.batchCount[base + threadIdx.x]);  // cross-CTA data after a grid barrier: always through L2
			unsigned int v = (raw + 31u) & ~31u;
			unsigned int incl = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
				if ((int)threadIdx.x >= o) incl += t;
			}
			s.batchOffset[base + threadIdx.x] = run + incl - v;
			unsigned int nz = __ballot_sync(0xffffffffu, raw != 0);
			if (nz) numBatches = base + 32 - __clz(nz);
			run += __shfl_sync(0xffffffffu, incl, 31);
		}
		if (threadIdx.x == 0)
		{
			s.batchOffset[MAX_BATCHES] = run;
			s.ctr[CTR_BATCHES] = (unsigned int)numBatches;
			s.ctr[CTR_COLOUR_ROUNDS] = (unsigned int)(round + 1);
		}
	}
	bar.sync();

	B3_PROBE("offsets");
	// ---- phase 3: contact -> constraint rows, written in batch order
	for (int k = tid; k < MAX_BATCHES * 32; k += stride)
	{
		const int bch = k >> 5;
		const unsigned int slot = __ldcg(&s.batchOffset[bch]) + __ldcg(&s.batchCount[bch]) + (unsigned int)(k & 31);
		if (slot < __ldcg(&s.batchOffset[bch + 1]))
		{
			float4* dw = reinterpret_cast<float4*>(&s.constraints[slot]);
			dw[6] = mk4(0, 0, 0, 0);
			dw[9] = mk4(0, 0, 0, 0);
			int4 tail;
			tail.x = -1;  // bodyA = 0xffffffff marks a padding slot
			tail.y = -1;
			tail.z = -1;
			tail.w = 0;
			reinterpret_cast<int4*>(dw)[10] = tail;
		}
	}
	for (int base = tid - lane; base < nContacts; base += stride)
	{
		const int c = base + lane;
		const int colour = c < nContacts ? __ldcg(&s.contactColour[c]) : -1;
		const unsigned int rank = warpCountByKey(s.batchCursor, colour, lane);
		if (colour < 0) continue;
		buildConstraint(s, &s.contacts[c], colour, &s.constraints[__ldcg(&s.batchOffset[colour]) + rank]);
	}
	B3_PROBE("built(block0 thread0 only)");
}

// ---------------------------------------------------------------- iterations
struct IterArgs
{
	b3b200_constraint4* constraints;
	const unsigned int* ctr;
	const float4* pose;
	float4* vel;
	const b3b200_inertia* inertias;
	const unsigned int* batchOffset;
	unsigned int* bar;
	int iterations;
	const unsigned long long* bodyMask;  // colours in use per body (from the setup kernel)
	unsigned int* seq;                   // per-body progress counter (dataflow kernel)
};

// Iteration arithmetic: explicit FMAs.  This file is built with --fmad=false, so the only fused operations are the
// ones written here and every kernel that inlines these helpers (barrier kernel, its one-CTA tail, dataflow kernel)
// produces the same bits for the same Gauss-Seidel order.  (The oracle evaluates the reference's unfused expressions;
// the bar for velocities is 1e-4 relative.)
B3_D float fdot3(const float4& a, const float4& b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, a.x * b.x)); }
B3_D float4 fcross3(const float4& a, const float4& b)
{
	return mk4(__fmaf_rn(a.y, b.z, -(a.z * b.y)), __fmaf_rn(a.z, b.x, -(a.x * b.z)), __fmaf_rn(a.x, b.y, -(a.y * b.x)));
}
B3_D float4 fmatRowMul(const float4& r0, const float4& r1, const float4& r2, const float4& v) { return mk4(fdot3(r0, v), fdot3(r1, v), fdot3(r2, v)); }
// v + (dir * k) * s
B3_D float4 faddScaled(const float4& v, const float4& dir, float k, float s)
{
	return mk4(__fmaf_rn(dir.x * k, s, v.x), __fmaf_rn(dir.y * k, s, v.y), __fmaf_rn(dir.z * k, s, v.z));
}
B3_D float4 faddScaled1(const float4& v, const float4& dir, float s) { return mk4(__fmaf_rn(dir.x, s, v.x), __fmaf_rn(dir.y, s, v.y), __fmaf_rn(dir.z, s, v.z)); }
B3_D float fcalcRelVel(const float4& l0, const float4& l1, const float4& a0, const float4& a1, const float4& linVel0, const float4& angVel0,
					   const float4& linVel1, const float4& angVel1)
{
	return fdot3(l0, linVel0) + fdot3(a0, angVel0) + fdot3(l1, linVel1) + fdot3(a1, angVel1);
}

struct RowData
{
	float4 lin, wp0, wp1, wp2, wp3, center, jac, bias, applied, fr;
	float4 posA, posB;
	float4 ia0, ia1, ia2, ib0, ib1, ib2;
	int aIdx, bIdx;
};

template <int PHASE>
B3_D void loadRow(const IterArgs& s, const b3b200_constraint4* __restrict__ cs, RowData& r)
{
	const float4* cw = reinterpret_cast<const float4*>(cs);
	const int4 tail = reinterpret_cast<const int4*>(cs)[10];
	r.aIdx = tail.x;
	r.bIdx = tail.y;
	r.lin = cw[0];
	r.applied = cw[8];
	if (PHASE == 0)
	{
		r.wp0 = cw[1];
		r.wp1 = cw[2];
		r.wp2 = cw[3];
		r.wp3 = cw[4];
		r.jac = cw[6];
		r.bias = cw[7];
	}
	else
	{
		r.center = cw[5];
		r.fr = cw[9];
	}
	if (r.aIdx < 0) return;  // padding slot
	r.posA = s.pose[2 * r.aIdx];
	r.posB = s.pose[2 * r.bIdx];
	const float4* IA = reinterpret_cast<const float4*>(&s.inertias[r.aIdx].invInertiaWorld);
	const float4* IB = reinterpret_cast<const float4*>(&s.inertias[r.bIdx].invInertiaWorld);
	r.ia0 = __ldg(IA);
	r.ia1 = __ldg(IA + 1);
	r.ia2 = __ldg(IA + 2);
	r.ib0 = __ldg(IB);
	r.ib1 = __ldg(IB + 1);
	r.ib2 = __ldg(IB + 2);
}

// solveContact<false> (b3Solver.cpp:187-266) on preloaded row data; explicit FMAs (see fdot3)
B3_D void solveNormalPre(const IterArgs& s, b3b200_constraint4* __restrict__ cs, const RowData& r)
{
	const int aIdx = r.aIdx, bIdx = r.bIdx;
	if (aIdx < 0) return;
	const float invMassA = r.posA.w, invMassB = r.posB.w;
	float4 linVelA = __ldcg(&s.vel[2 * aIdx]), angVelA = __ldcg(&s.vel[2 * aIdx + 1]);
	float4 linVelB = __ldcg(&s.vel[2 * bIdx]), angVelB = __ldcg(&s.vel[2 * bIdx + 1]);
	const float4 n = mk4(r.lin.x, r.lin.y, r.lin.z);
	const float4 nn = neg3(n);
	const float jacv[4] = {r.jac.x, r.jac.y, r.jac.z, r.jac.w};
	const float bv[4] = {r.bias.x, r.bias.y, r.bias.z, r.bias.w};
	float ap[4] = {r.applied.x, r.applied.y, r.applied.z, r.applied.w};
#pragma unroll
	for (int ic = 0; ic < 4; ic++)
	{
		if (jacv[ic] == 0.f) continue;
		const float4 wp = ic == 0 ? r.wp0 : (ic == 1 ? r.wp1 : (ic == 2 ? r.wp2 : r.wp3));
		float4 r0 = sub3(wp, r.posA), r1 = sub3(wp, r.posB);
		float4 angular0 = fcross3(r0, n);
		float4 angular1 = neg3(fcross3(r1, n));
		float rambdaDt = fcalcRelVel(n, nn, angular0, angular1, linVelA, angVelA, linVelB, angVelB) + bv[ic];
		rambdaDt *= jacv[ic];
		{
			float prevSum = ap[ic];
			float updated = prevSum;
			updated += rambdaDt;
			updated = fmaxf(updated, 0.f);
			updated = fminf(updated, FLT_MAX);
			rambdaDt = updated - prevSum;
			ap[ic] = updated;
		}
		linVelA = faddScaled(linVelA, n, invMassA, rambdaDt);
		angVelA = faddScaled1(angVelA, fmatRowMul(r.ia0, r.ia1, r.ia2, angular0), rambdaDt);
		linVelB = faddScaled(linVelB, nn, invMassB, rambdaDt);
		angVelB = faddScaled1(angVelB, fmatRowMul(r.ib0, r.ib1, r.ib2, angular1), rambdaDt);
	}
	reinterpret_cast<float4*>(cs)[8] = mk4(ap[0], ap[1], ap[2], ap[3]);
	if (invMassA != 0.f)
	{
		__stcg(&s.vel[2 * aIdx], linVelA);
		__stcg(&s.vel[2 * aIdx + 1], angVelA);
	}
	if (invMassB != 0.f)
	{
		__stcg(&s.vel[2 * bIdx], linVelB);
		__stcg(&s.vel[2 * bIdx + 1], angVelB);
	}
}

// solveFriction (b3Solver.cpp:268-329) on preloaded row data; explicit FMAs (see fdot3)
B3_D void solveFrictionPre(const IterArgs& s, b3b200_constraint4* __restrict__ cs, const RowData& r)
{
	const float4 fr = r.fr;
	if (fr.x == 0.f && fr.x == 0.f) return;
	const int aIdx = r.aIdx, bIdx = r.bIdx;
	const float4 posA = r.posA, posB = r.posB;
	const float invMassA = posA.w, invMassB = posB.w;
	float4 linVelA = __ldcg(&s.vel[2 * aIdx]), angVelA = __ldcg(&s.vel[2 * aIdx + 1]);
	float4 linVelB = __ldcg(&s.vel[2 * bIdx]), angVelB = __ldcg(&s.vel[2 * bIdx + 1]);
	float sum = 0.f;
	sum += r.applied.x;
	sum += r.applied.y;
	sum += r.applied.z;
	sum += r.applied.w;
	const float frictionCoeff = 0.7f;
	const float maxR = frictionCoeff * sum;
	const float minR = -maxR;
	const float4 n = neg3(mk4(r.lin.x, r.lin.y, r.lin.z));
	float4 tangent[2];
	planeSpace1(n, tangent[0], tangent[1]);
	const float4 r0 = sub3(r.center, posA), r1 = sub3(r.center, posB);
	float fj[2] = {fr.x, fr.y};
	float fa[2] = {fr.z, fr.w};
#pragma unroll
	for (int i = 0; i < 2; i++)
	{
		const float4 t = tangent[i];
		const float4 angular0 = fcross3(r0, t);
		const float4 angular1 = neg3(fcross3(r1, t));
		float rambdaDt = fcalcRelVel(t, neg3(t), angular0, angular1, linVelA, angVelA, linVelB, angVelB);
		rambdaDt *= fj[i];
		{
			float prevSum = fa[i];
			float updated = prevSum;
			updated += rambdaDt;
			updated = fmaxf(updated, minR);
			updated = fminf(updated, maxR);
			rambdaDt = updated - prevSum;
			fa[i] = updated;
		}
		linVelA = faddScaled(linVelA, t, invMassA, rambdaDt);
		angVelA = faddScaled1(angVelA, fmatRowMul(r.ia0, r.ia1, r.ia2, angular0), rambdaDt);
		linVelB = faddScaled(linVelB, neg3(t), invMassB, rambdaDt);
		angVelB = faddScaled1(angVelB, fmatRowMul(r.ib0, r.ib1, r.ib2, angular1), rambdaDt);
	}
	{
		// angular damping for point constraint (b3Solver.cpp:317-328)
		float4 ab = normalized3(sub3(posB, posA));
		float4 ac = normalized3(sub3(r.center, posA));
		if (fdot3(ab, ac) > 0.95f || (invMassA == 0.f || invMassB == 0.f))
		{
			float angNA = fdot3(n, angVelA);
			float angNB = fdot3(n, angVelB);
			angVelA = faddScaled1(angVelA, n, -(angNA * 0.1f));
			angVelB = faddScaled1(angVelB, n, -(angNB * 0.1f));
		}
	}
	reinterpret_cast<float4*>(cs)[9] = mk4(fj[0], fj[1], fa[0], fa[1]);
	if (invMassA != 0.f)
	{
		__stcg(&s.vel[2 * aIdx], linVelA);
		__stcg(&s.vel[2 * aIdx + 1], angVelA);
	}
	if (invMassB != 0.f)
	{
		__stcg(&s.vel[2 * bIdx], linVelB);
		__stcg(&s.vel[2 * bIdx + 1], angVelB);
	}
}

template <int PHASE>
B3_D void iteratePhase(const IterArgs& s, GridBarrier& bar, int numBatches, int tailStart, int stride, int firstOffset)
{
	RowData pre;
	bool havePre = false;
	if (tailStart > 0)
	{
		const int i0 = (int)s.batchOffset[0] + firstOffset;
		if (i0 < (int)s.batchOffset[1])
		{
			loadRow<PHASE>(s, &s.constraints[i0], pre);
			havePre = true;
		}
	}
	for (int iter = 0; iter < s.iterations; iter++)
	{
		for (int b = 0; b < tailStart; b++)
		{
			// warp-rows are dealt round-robin to the CTAs so that a small batch still uses every SM
			const int begin = (int)s.batchOffset[b], end = (int)s.batchOffset[b + 1];
			for (int i = begin + firstOffset; i < end; i += stride)
			{
				if (!havePre) loadRow<PHASE>(s, &s.constraints[i], pre);
				havePre = false;
				if (PHASE == 0)
					solveNormalPre(s, &s.constraints[i], pre);
				else
					solveFrictionPre(s, &s.constraints[i], pre);
			}
			bar.arrive();
			// while the other CTAs arrive: fetch this thread's first row of the next grid-wide batch
			{
				int nb = b + 1;
				bool more = true;
				if (nb == tailStart)
				{
					nb = 0;
					more = tailStart == numBatches && iter + 1 < s.iterations;  // with a tail, batch 0 is prefetched after it
				}
				if (more)
				{
					const int i2 = (int)s.batchOffset[nb] + firstOffset;
					if (i2 < (int)s.batchOffset[nb + 1])
					{
						loadRow<PHASE>(s, &s.constraints[i2], pre);
						havePre = true;
					}
				}
			}
			bar.wait();
		}
		if (tailStart < numBatches)
		{
			// The colouring leaves a long tail of small batches (a few hundred rows each).  A grid barrier costs more
			// than solving one of them, so CTA 0 runs the whole tail alone, batch after batch in the same order,
			// separated by __syncthreads(); everybody meets at ONE grid barrier afterwards.
			if (blockIdx.x == 0)
			{
				for (int b = tailStart; b < numBatches; b++)
				{
					const int begin = (int)s.batchOffset[b], end = (int)s.batchOffset[b + 1];
					for (int i = begin + (int)threadIdx.x; i < end; i += (int)blockDim.x)
					{
						RowData r;
						loadRow<PHASE>(s, &s.constraints[i], r);
						if (PHASE == 0)
							solveNormalPre(s, &s.constraints[i], r);
						else
							solveFrictionPre(s, &s.constraints[i], r);
					}
					__syncthreads();
				}
			}
			bar.arrive();
			if (tailStart > 0 && iter + 1 < s.iterations)
			{
				const int i2 = (int)s.batchOffset[0] + firstOffset;
				if (i2 < (int)s.batchOffset[1])
				{
					loadRow<PHASE>(s, &s.constraints[i2], pre);
					havePre = true;
				}
			}
			bar.wait();
		}
	}
}

template <int THREADS, int MIN_BLOCKS>
B3_D void solverIterateBody(const IterArgs& s)
{
	GridBarrier bar;
	bar.init(s.bar, gridDim.x);
	const int stride = gridDim.x * blockDim.x;
	const int numBatches = (int)s.ctr[CTR_BATCHES];
	if (numBatches == 0) return;
	const int firstOffset = (((threadIdx.x >> 5) * (int)gridDim.x + (int)blockIdx.x) << 5) + (threadIdx.x & 31);
	// tail = the trailing run of batches with at most TAIL_ROWS rows each (solved by CTA 0 alone, see iteratePhase)
	int tailStart = numBatches;
	while (tailStart > 0 && (int)(s.batchOffset[tailStart] - s.batchOffset[tailStart - 1]) <= (int)blockDim.x) tailStart--;
	iteratePhase<0>(s, bar, numBatches, tailStart, stride, firstOffset);
	iteratePhase<1>(s, bar, numBatches, tailStart, stride, firstOffset);
}

__global__ void __launch_bounds__(ITER_THREADS) solverIterateKernel(IterArgs s) { solverIterateBody<ITER_THREADS, 1>(s); }
// (measured: 2 CTAs per SM of 384 / 512 threads, i.e. 80 / 64 registers with part of the prefetched row spilled, take 2.74 /
// 3.17 ms against 2.39 ms for this 128-register, one-CTA-per-SM version on the bench scene)

// ---------------------------------------------------------------- dataflow iterations
// Same Gauss-Seidel order as solverIterateKernel, no grid-wide barriers.  Within a body, the
// constraints touching it are totally ordered by (round, batch); constraints that share no
// dynamic body commute exactly.  So each constraint only has to wait for ITS two bodies:
// every dynamic body carries a progress counter seq[b]; constraint c with colour k is the
// rank(c,b) = popc(mask[b] & below(k))-th user of body b in every round, and may run in round r
// when seq[b] == r * deg(b) + rank(c,b) for both bodies.  It then solves, publishes the new
// velocities and bumps both counters (release/acquire at gpu scope).  Threads own constraints
// i = tid, tid+T, ... of the batch-sorted array and visit them in ascending order each round,
// which is consistent with the global (round, batch) order, so the earliest unfinished
// constraint is always runnable: no deadlock as long as all threads are resident
// (cooperative launch).  Ready lanes solve inside the polling loop, so lanes of one warp never
// wait for each other at a reconvergence point.
constexpr int DF_THREADS = 256;

B3_D unsigned int ldRelaxed(const unsigned int* p)
{
	unsigned int v;
	asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
B3_D void stRelaxed(unsigned int* p, unsigned int v) { asm volatile("st.relaxed.gpu.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

B3_D void rankAndDegree(const unsigned long long* __restrict__ mask, int body, int colour, bool& dyn, unsigned int& rank, unsigned int& deg)
{
	const unsigned long long m0 = __ldg(&mask[2 * body]), m1 = __ldg(&mask[2 * body + 1]);
	const unsigned long long bit = 1ull << (colour & 63);
	dyn = ((colour < 64 ? m0 : m1) & bit) != 0ull;
	deg = __popcll(m0) + __popcll(m1);
	rank = colour < 64 ? __popcll(m0 & (bit - 1ull)) : __popcll(m0) + __popcll(m1 & (bit - 1ull));
}

// One THREAD owns the slots i = tid, tid + T, ... of the batch-sorted array and walks them in ascending order every round.
// A lane whose two bodies have reached its row's turn solves it at once, inside the polling loop; the other lanes of the
// warp keep polling (independent thread scheduling), nobody waits at a reconvergence point.  The row, the positions and
// the inertias are fetched when the thread moves on to the slot (only this thread ever writes the row), so when the
// counters match only the velocities are still to be loaded.
template <int PHASE>
B3_D void dataflowRounds(const IterArgs& s, int tid, int T, int nSlots, int roundBegin, int roundEnd)
{
	int round = roundBegin, i = tid;
	bool have = false, valid = false, dynA = false, dynB = false;
	unsigned int expA = 0, expB = 0;
	RowData pre;
	while (round < roundEnd)
	{
		if (!have)
		{
			loadRow<PHASE>(s, &s.constraints[i], pre);
			valid = pre.aIdx >= 0;
			if (valid)
			{
				const int colour = reinterpret_cast<const int4*>(&s.constraints[i])[10].z;
				unsigned int rank, deg;
				rankAndDegree(s.bodyMask, pre.aIdx, colour, dynA, rank, deg);
				expA = (unsigned int)round * deg + rank;
				rankAndDegree(s.bodyMask, pre.bIdx, colour, dynB, rank, deg);
				expB = (unsigned int)round * deg + rank;
			}
			have = true;
		}
		bool ready = true;
		if (valid)
		{
			if (dynA) ready = ldRelaxed(&s.seq[pre.aIdx]) == expA;
			if (ready && dynB) ready = ldRelaxed(&s.seq[pre.bIdx]) == expB;
		}
		if (ready)
		{
			if (valid)
			{
				// No acquire fence here: it would invalidate the whole L1 (CCTL.IVALL) once per row.  Everything another
				// thread writes (velocities, counters) is read with strong gpu-scope loads served by L2, and those loads
				// are only issued once the branch on the counter values has resolved.
				if (PHASE == 0)
					solveNormalPre(s, &s.constraints[i], pre);
				else
					solveFrictionPre(s, &s.constraints[i], pre);
				if (dynA || dynB) __threadfence();  // the velocity stores are performed before either counter moves
				if (dynA) stRelaxed(&s.seq[pre.aIdx], expA + 1u);
				if (dynB) stRelaxed(&s.seq[pre.bIdx], expB + 1u);
			}
			have = false;
			i += T;
			if (i >= nSlots)
			{
				i = tid;
				round++;
			}
		}
	}
}

__global__ void __launch_bounds__(DF_THREADS) solverIterateDataflowKernel(IterArgs s)
{
	const int T = gridDim.x * blockDim.x;
	// consecutive lanes own consecutive slots; warps are dealt round-robin to the CTAs like in the barrier kernel
	const int tid = (((threadIdx.x >> 5) * (int)gridDim.x + (int)blockIdx.x) << 5) + (threadIdx.x & 31);
	const int nSlots = (int)s.batchOffset[MAX_BATCHES];  // batches are padded to multiples of 32
	if (tid >= nSlots) return;
	dataflowRounds<0>(s, tid, T, nSlots, 0, s.iterations);
	dataflowRounds<1>(s, tid, T, nSlots, s.iterations, 2 * s.iterations);
}

static int coopLaunch(World* w, const void* fn, void* argStruct, int threads = SOLVER_THREADS)
{
	int perSm = 0;
	B3_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, threads, 0));
	if (perSm < 1)
	{
		setLastError("solver kernel does not fit on an SM");
		return B3B200_ERR_CUDA;
	}
	if (perSm > 2) perSm = 2;
	dim3 grid(w->smCount * perSm), block(threads);
	void* args[] = {argStruct};
	B3_CUDA_CHECK(cudaMemsetAsync(w->dGridBarrier.ptr, 0, sizeof(unsigned int) * 4, w->stream));
	B3_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, grid, block, args, 0, w->stream));
	g_launchCount++;
	return 0;
}

int launchSolverSetup(World* w)
{
	SetupArgs s;
	s.contacts = w->dContacts.ptr;
	s.ctr = w->dCounters.ptr;
	s.pose = w->dPose.ptr;
	s.vel = w->dVel.ptr;
	s.inertias = w->dInertias.ptr;
	s.constraints = w->dConstraints.ptr;
	s.bodyMask = w->dBodyMask.ptr;
	s.bodyPrio = reinterpret_cast<unsigned long long*>(w->dBodyPrio.ptr);
	s.contactColour = w->dContactColour.ptr;
	s.batchCount = w->dBatchCount.ptr;
	s.batchOffset = w->dBatchOffset.ptr;
	s.batchCursor = w->dBatchCursor.ptr;
	s.remaining = w->dBodyCount.ptr;  // MAX_ROUNDS words, see World::init
	s.colourList = w->dColourList.ptr;
	s.colourListStride = (int)(w->dColourList.cap / 2);
	s.bar = w->dGridBarrier.ptr;
	s.numBodies = w->numBodies;
	s.staticIdx = w->static0Index;
	s.colouring = w->solverColouring;
	s.dt = 1.f / 60.f;  // the reference solver ignores deltaTime (b3GpuPgsContactSolver.cpp:672)
	s.positionDrift = 0.005f;
	s.positionConstraintCoeff = 0.2f;
	return coopLaunch(w, (const void*)solverSetupKernel, &s);
}

int launchSolverIterate(World* w)
{
	IterArgs s;
	s.constraints = w->dConstraints.ptr;
	s.ctr = w->dCounters.ptr;
	s.pose = w->dPose.ptr;
	s.vel = w->dVel.ptr;
	s.inertias = w->dInertias.ptr;
	s.batchOffset = w->dBatchOffset.ptr;
	s.bar = w->dGridBarrier.ptr;
	s.iterations = w->solverIterations;
	s.bodyMask = w->dBodyMask.ptr;
	s.seq = w->dBodyCount.ptr;
	w->soaDirty = true;
	if (w->solverDataflow)
	{
		// all threads must be co-resident: grid = SMs x occupancy, cooperative launch
		int perSm = 0;
		B3_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, (const void*)solverIterateDataflowKernel, DF_THREADS, 0));
		if (perSm < 1)
		{
			setLastError("solver kernel does not fit on an SM");
			return B3B200_ERR_CUDA;
		}
		B3_CUDA_CHECK(cudaMemsetAsync(w->dBodyCount.ptr, 0, sizeof(unsigned int) * (size_t)std::max(w->numBodies, 1), w->stream));
		dim3 grid(w->smCount * perSm), block(DF_THREADS);
		void* args[] = {&s};
		B3_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)solverIterateDataflowKernel, grid, block, args, 0, w->stream));
		g_launchCount++;
		return 0;
	}
	return coopLaunch(w, (const void*)solverIterateKernel, &s, ITER_THREADS);
}

}  // namespace b3b200
