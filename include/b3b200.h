/*
 * b3b200.h -- C ABI of the B200-native GPU rigid-body step.
 *
 * This is the drop-in boundary underneath the reference's C++ class surface
 * (src/Bullet3OpenCL).  The reference has no FFI layer of its own: callers use
 * the classes directly (examples/OpenCL/rigidbody/GpuRigidBodyDemo.cpp:129-150,
 * examples/OpenCL/broadphase/PairBench.cpp:208-379).  The same-named C++
 * classes under bullet3_b200/csrc/host/ forward to the entry points below, and
 * each entry point cites the reference method it replaces.
 *
 * Conventions: every function returns an int.  Functions that create an index
 * return the index (>= 0) or -1 on failure, like the reference's register*
 * methods (b3GpuNarrowPhase.cpp:144-157, 821-825).  All other functions return
 * 0 on success and a negative B3B200_ERR_* code on failure.  Nothing throws or
 * aborts; the text of the last failure is available from b3b200_last_error().
 * All pointers are HOST pointers unless the name says "device".
 * A world is bound to one device + one stream; worlds are independent of each
 * other (re-entrant across worlds, not thread-safe per world).
 */
#ifndef B3B200_H
#define B3B200_H

#include "b3b200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B3B200_OK 0
#define B3B200_ERR_INVALID -1  /* bad argument / bad index */
#define B3B200_ERR_CAPACITY -2 /* a b3Config capacity would be exceeded */
#define B3B200_ERR_CUDA -3     /* CUDA runtime failure (see b3b200_last_error) */
#define B3B200_ERR_STATE -4    /* call order violated (e.g. step before upload) */

typedef struct b3b200_world b3b200_world;
typedef struct b3b200_broadphase b3b200_broadphase;

/* broadphase kinds: b3GpuSapBroadphase (b3GpuSapBroadphase.h:14-141) and
 * b3GpuGridBroadphase (b3GpuGridBroadphase.h:7-78). */
#define B3B200_BP_SAP 0
#define B3B200_BP_GRID 1
/* contact solver kinds: b3GpuPgsContactSolver (default, b3GpuRigidBodyPipeline.cpp:449)
 * and b3GpuJacobiContactSolver (gUseJacobi, :389-447). */
#define B3B200_SOLVER_PGS 0
#define B3B200_SOLVER_JACOBI 1

const char* b3b200_last_error(void);
int b3b200_version(void);
/* number of kernel launches issued by this library since load (all worlds) */
long long b3b200_launch_count(void);

/* b3Config::b3Config() defaults (b3Config.h:19-36) */
int b3b200_config_default(b3b200_config* cfg);

/* ------------------------------------------------------------------ world */
/* replaces: new b3GpuNarrowPhase + new b3Gpu{Sap,Grid}Broadphase + new b3GpuRigidBodyPipeline
 * (GpuRigidBodyDemo.cpp:129-146; b3GpuRigidBodyPipeline.cpp:65-114).
 * `stream` is a cudaStream_t (0 = a private non-blocking stream is created).
 * device == -1 creates a HOST-ONLY world: shapes and bodies can be registered and the
 * tables read back (b3b200_get_table), every call that needs the GPU fails with
 * B3B200_ERR_STATE.  There is no CPU implementation of the step. */
int b3b200_create(const b3b200_config* cfg, int device, void* stream, b3b200_world** out);
int b3b200_destroy(b3b200_world* w);
/* b3GpuRigidBodyPipeline::reset + b3GpuNarrowPhase::reset (b3GpuRigidBodyPipeline.cpp:141-150) */
int b3b200_reset(b3b200_world* w);

/* ---- shapes: b3GpuNarrowPhase::register*Shape (b3GpuNarrowPhase.cpp:159-668) ---- */
/* registerConvexHullShape(b3ConvexUtility*) :321-368 -- a ready polyhedron.
 * `poly` supplies localCenter/extents/mC/mE/radius; its offsets/counts are ignored. */
int b3b200_register_convex(b3b200_world* w,
						   const b3b200_float4* vertices, int numVertices,
						   const b3b200_face* faces, int numFaces,
						   const int* indices, int numIndices,
						   const b3b200_float4* uniqueEdges, int numUniqueEdges,
						   const b3b200_convex_polyhedron* poly);
/* registerConvexHullShape(const float*, stride, n, scaling) :298-319 -- builds the hull */
int b3b200_register_convex_points(b3b200_world* w, const float* vertices, int strideInBytes,
								  int numVertices, const float* scaling3);
/* registerPlaneShape :196-231, registerSphereShape :159-194 */
int b3b200_register_plane(b3b200_world* w, const float* normal3, float planeConstant);
int b3b200_register_sphere(b3b200_world* w, float radius);
/* registerCompoundShape :370-474 -- children reference convex collidables via shapeIndex */
int b3b200_register_compound(b3b200_world* w, const b3b200_child_shape* children, int numChildren);
/* registerConcaveMesh (b3GpuNarrowPhase.cpp:521-605) + registerConcaveMeshShape (:607-668): triangles become the faces of
 * one convex-table entry; this build adds its own float AABB tree instead of the quantized b3OptimizedBvh */
int b3b200_register_concave(b3b200_world* w, const float* vertices, int numVertices,
							const int* triIndices, int numIndices, const float* scaling3);

/* ---- bodies ---- */
/* b3GpuRigidBodyPipeline::registerPhysicsInstance (b3GpuRigidBodyPipeline.cpp:603-669):
 * world AABB with margin 0.01, registerRigidBody, createProxy / createLargeProxy. */
int b3b200_register_instance(b3b200_world* w, float mass, const float* position,
							 const float* orientation, int collidableIndex, int userIndex);
/* b3GpuNarrowPhase::registerRigidBody (b3GpuNarrowPhase.cpp:816-908) with a caller-supplied world AABB (used for the
 * box-approximated inertia, :876-897).  Unlike the reference this also creates the broadphase proxy, because the world
 * keeps AABB index == body index. */
int b3b200_register_body(b3b200_world* w, int collidableIndex, float mass, const float* position, const float* orientation,
						 const float* aabbMin3, const float* aabbMax3);
/* the same for n instances in one call (positions/orientations: n x 4 floats); returns the first body index */
/* The MPR stage of the reference's GPU narrowphase as a stand-alone kernel: mprPenetrationKernel (kernels/mpr.cl:14-89, launched
 * by b3ConvexHullContact.cpp:2817-2850 when useMprGpu is set) = b3MprPenetration (shared/b3MprPenetration.h:825-888) per convex x
 * convex pair; a penetrating pair gets pairs[i].z = its contact, a ONE-point contact (portal position, normal = -direction,
 * depth = -distance) appended at *numContacts, hasSepAxis[i] = 1 and sepNormals[i] = -direction.  All arrays are HOST arrays
 * in the layout of the reference kernel's buffers (this is the call the reference's serialized test launches replay,
 * test/OpenCL/AllBullet3Kernels/testExecuteBullet3NarrowphaseKernels.cpp:382-420). */
int b3b200_mpr_penetration(int device, b3b200_int4* pairs, int numPairs, const b3b200_rigid_body* bodies, int numBodies,
						   const b3b200_collidable* collidables, int numCollidables, const b3b200_convex_polyhedron* convex, int numConvex,
						   const b3b200_float4* vertices, int numVertices, b3b200_float4* sepNormals, int* hasSepAxis,
						   b3b200_contact4* contactsOut, int contactCapacity, int* numContacts, b3b200_mpr_result* results);
/* Pipelined stepping from / to HOST memory: upload `hostIn` (all bodies; NULL = keep stepping the state on the device), step,
 * download the stepped bodies into `hostOut` (NULL = no download) -- all asynchronously: the upload of call c overlaps the
 * step of call c - 1 and the download overlaps the step of call c + 1 (two staging slots, separate copy streams).  The host
 * buffers must be page-locked (cudaHostAlloc / cudaHostRegister) for the overlap and must stay untouched until
 * b3b200_step_host_wait returns.  The blocking reference loop it replaces: writeAllBodiesToGpu -> stepSimulation ->
 * readbackAllBodiesToCpu (b3GpuNarrowPhase.cpp:1020-1040, b3GpuRigidBodyPipeline.cpp:221-463). */
int b3b200_step_host_async(b3b200_world* w, float dt, const b3b200_rigid_body* hostIn, b3b200_rigid_body* hostOut, int numBodies);
int b3b200_step_host_wait(b3b200_world* w);
/* Batched independent worlds (no counterpart in the reference, which steps one world per b3GpuRigidBodyPipeline):
 * bodies registered after b3b200_set_current_world(w, k) belong to world k (default 0).  Bodies of different worlds never
 * collide, wherever they stand; every world sees exactly the pairs and contacts it would see alone.  One broadphase pass,
 * one narrowphase, one solve for all worlds: the world id is part of the grid broadphase's cell key and of the solver's block
 * order (blocks hold whole worlds when the worlds have equally many dynamic bodies: no cross-block contacts, no grid
 * barriers).  Static bodies belong to a world like any other body.  With the SAP broadphase the result is the same but the
 * sweep sees all worlds interleaved (slow): use the grid. */
int b3b200_set_current_world(b3b200_world* w, int worldIndex);
int b3b200_num_worlds(b3b200_world* w);
int b3b200_get_body_worlds(b3b200_world* w, int* dst, int n);
int b3b200_register_instances(b3b200_world* w, int n, const float* masses, const float* positions4,
							  const float* orientations4, const int* collidableIndices);
/* writeAllInstancesToGpu + writeAllBodiesToGpu + writeAabbsToGpu (GpuRigidBodyDemo.cpp:148-150).  Bodies that are already on the
 * device keep their device state (poses, velocities): only the shape tables and the bodies registered since the last
 * upload are sent, like the reference's copyFromHostPointer(&body, 1, bodyIndex) per new body. */
int b3b200_upload(b3b200_world* w);

/* ---- joints: b3GpuRigidBodyPipeline::createPoint2PointConstraint / createFixedConstraint / removeConstraintByUid /
 * copyConstraintsToHost (b3GpuRigidBodyPipeline.cpp:158-218, 593-596).  They take effect at the next step; like in the
 * reference the joints are solved before the contacts (4 iterations, ERP 0.2, dt 1/60), a joint whose applied impulse
 * reaches its breaking threshold is disabled (flags = 0).  create_* return the uid (>= 0) or -1. */
int b3b200_create_p2p_constraint(b3b200_world* w, int bodyA, int bodyB, const float* pivotInA3, const float* pivotInB3, float breakingThreshold);
int b3b200_create_fixed_constraint(b3b200_world* w, int bodyA, int bodyB, const float* pivotInA3, const float* pivotInB3, const float* relTargetAB4,
								   float breakingThreshold);
int b3b200_remove_constraint(b3b200_world* w, int uid);
int b3b200_num_constraints(b3b200_world* w);
int b3b200_get_joints(b3b200_world* w, b3b200_generic_constraint* dst, int capacity, int* count);

/* ---- b3GpuRigidBodyPipeline::castRays (b3GpuRigidBodyPipeline.cpp:671-680 -> b3GpuRaycast::castRays,
 * src/Bullet3OpenCL/Raycast/b3GpuRaycast.cpp:249-373; result semantics of its host twin castRaysHost :160-246): for
 * every ray the closest hit with hitFraction < the value the caller stored in hits[i].hitFraction (normally 1);
 * convex hulls and spheres are tested, other shape types are ignored like in the reference; the hit normal of a hull
 * is the face normal in the hull's LOCAL frame (reference quirk).  rays / hits are HOST arrays of numRays entries;
 * hits of rays that hit nothing are left untouched. */
int b3b200_cast_rays(b3b200_world* w, const b3b200_ray_info* rays, int numRays, b3b200_ray_hit* hits);
/* how cast_rays culls: -1 (default) = per call: a linear BVH over the Morton-sorted bodies, rebuilt for the call (the role of
 * b3GpuParallelLinearBvh), when there are >= 64 rays and >= 4096 dynamic bodies, else brute force over the world AABBs;
 * 0 = always brute force; 1 = always the tree.  The results are identical. */
int b3b200_set_ray_accel(b3b200_world* w, int mode);
/* b3GpuRigidBodyPipeline::setGravity (b3GpuRigidBodyPipeline.cpp:562-565) */
int b3b200_set_gravity(b3b200_world* w, const float* gravity3);
int b3b200_set_solver(b3b200_world* w, int kind, int iterations);
int b3b200_set_broadphase(b3b200_world* w, int kind);
/* how the PGS solver assigns contacts to batches (the reference: b3Solver::batchContacts / sortConstraintByBatch3; any
 * assignment in which no two contacts of a batch share a dynamic body gives a valid Gauss-Seidel order).  Two levels, like the
 * reference's cells: the dynamic bodies are cut into spatial blocks (one per SM); contacts inside a block are coloured by the
 * block's CTA in shared memory, contacts between blocks globally.  b3Contact4::m_batchIdx = position in the solve order.
 * 1 (default) = every contact takes the lowest colour free on both bodies with atomics (one pass; the colours depend on how
 * the races resolve); 0 = priority rounds (reproducible for a given contact array; the cross contacts in one CTA: slow
 * on large scenes). */
int b3b200_set_colouring(b3b200_world* w, int mode);
/* 1 (default): b3b200_step / b3b200_step_n replay the whole step as ONE captured CUDA graph (the reference issues each of its
 * ~40 kernels with a clFinish in between, b3GpuRigidBodyPipeline.cpp:221-463); 0: kernel by kernel.  Same kernels, same
 * results; worlds with joints, the Jacobi solver or stage timing enabled always step kernel by kernel. */
int b3b200_set_step_graphs(b3b200_world* w, int on);
/* clip window of the convex-convex clipper: the reference kernels use
 * (-1e30, 0.02) (satClipHullContacts.cl:916-917), the shared CPU header (-1, 0)
 * (b3ContactConvexConvexSAT.h:320-321).  Default = the kernel constants. */
int b3b200_set_contact_clip(b3b200_world* w, float minDist, float maxDist);
int b3b200_set_angular_damping(b3b200_world* w, float damping);

/* overwrite / read the body state (b3RigidBodyData AoS, 80 B each) */
int b3b200_write_bodies(b3b200_world* w, const b3b200_rigid_body* src, int n);
/* b3GpuNarrowPhase::readbackAllBodiesToCpu + getBodiesCpu (b3GpuNarrowPhase.cpp:965-968, 676-679) */
int b3b200_readback_bodies(b3b200_world* w, b3b200_rigid_body* dst, int n);
int b3b200_readback_inertias(b3b200_world* w, b3b200_inertia* dst, int n);
/* one body: copyFromHostPointer(&body, 1, bodyIndex) / the matching read (b3GpuNarrowPhase::setObjectTransform, b3GpuNarrowPhase.cpp:946-963;
 * the pick-and-drag flow of GpuRigidBodyDemo.cpp:457-463) */
int b3b200_write_body(b3b200_world* w, int bodyIndex, const b3b200_rigid_body* src);
int b3b200_read_body(b3b200_world* w, int bodyIndex, b3b200_rigid_body* dst);
int b3b200_num_bodies(b3b200_world* w);

/* ---- the step: b3GpuRigidBodyPipeline::stepSimulation (b3GpuRigidBodyPipeline.cpp:221-463) ---- */
int b3b200_step(b3b200_world* w, float dt);
/* run `n` steps back to back with no host synchronisation in between */
int b3b200_step_n(b3b200_world* w, float dt, int n);
/* waits for the world's stream.  A capacity overrun of the steps since (pairs, contacts, batches, work items: the buffers were clamped,
 * b3b200_get_counters()[4] holds the flags) is reported here through b3b200_last_error(); the call itself still returns 0. */
int b3b200_synchronize(b3b200_world* w);

/* per-stage entry points (parity tests call these one at a time) */
int b3b200_update_aabbs(b3b200_world* w);      /* setupGpuAabbsFull :500-560 */
int b3b200_find_pairs(b3b200_world* w);        /* bp->calculateOverlappingPairs :260 */
int b3b200_compute_contacts(b3b200_world* w);  /* np->computeContacts :324 */
int b3b200_solve_contacts(b3b200_world* w);    /* m_solver2/3->solveContacts :389-460 */
int b3b200_solve_joints(b3b200_world* w);      /* m_gpuSolver->solveJoints :361-374 (b3GpuPgsConstraintSolver.cpp:927-942) */
int b3b200_solver_setup(b3b200_world* w);      /* colouring + contact->constraint only */
int b3b200_solver_iterate(b3b200_world* w);    /* the iteration loop only */
int b3b200_integrate(b3b200_world* w, float dt); /* integrate :465-498 */

/* results (device -> host copies; each synchronises the world's stream) */
int b3b200_get_aabbs(b3b200_world* w, b3b200_aabb* dst, int n);
int b3b200_get_pairs(b3b200_world* w, b3b200_int4* dst, int capacity, int* numPairs);
int b3b200_get_contacts(b3b200_world* w, b3b200_contact4* dst, int capacity, int* numContacts);
int b3b200_set_contacts(b3b200_world* w, const b3b200_contact4* src, int numContacts);
/* the solver rows as b3ContactConstraint4 records in solve order (sorted by batch; inside a batch the order is free);
 * batchOffsets has numBatches+1 entries.  (Converted on request from the solver's own 32-row tiles.) */
int b3b200_get_constraints(b3b200_world* w, b3b200_constraint4* dst, int capacity, int* numConstraints);
int b3b200_get_batches(b3b200_world* w, int* batchOffsets, int capacity, int* numBatches);
/* counters of the last step: [0]=pairs [1]=contacts [2]=batches [3]=colouring rounds
 * [4]=overflow flags [5]=raw compound child pairs [6]=raw (pair, triangle, child) items [7]=SAT work items */
int b3b200_get_counters(b3b200_world* w, int* dst8);
/* the raw device counters of the last step (n <= 24) (diagnostics; no reference counterpart): [0..7] as above except [7] =
 * uncoloured contacts, [8] = SAT work items, [9] = overlapping items, [10] = halo records, [11] = triangle items that
 * passed the quick reject, [12..14] = work cursors, [15] = trimesh pairs, [16] = small x small hull items (box-like pairs), [17] = the other small pairs, [18] = overlapping items passed on to the warp-per-item clip, [19] = triangle items with a larger hull B ([11] = those with a small hull B) */
int b3b200_get_work_counters(b3b200_world* w, int* dst, int n);
/* ms per stage of the last step when timing is enabled:
 * [0]=aabbs [1]=broadphase [2]=narrowphase [3]=solver setup [4]=solver iterations [5]=integrate [6]=total */
int b3b200_enable_stage_timing(b3b200_world* w, int enable);
int b3b200_stage_timings(b3b200_world* w, float* ms8);

/* device pointers with the reference AoS layouts: getBodyBuffer(), getAabbBufferWS(),
 * getOverlappingPairBuffer(), getContactsGpu() (b3GpuRigidBodyPipeline.cpp:577,
 * b3GpuBroadphaseInterface.h:29-31, b3GpuNarrowPhase.cpp:713-721) */
#define B3B200_BUF_BODIES 0
#define B3B200_BUF_AABBS 1
#define B3B200_BUF_PAIRS 2
#define B3B200_BUF_CONTACTS 3
#define B3B200_BUF_INERTIAS 4
int b3b200_device_buffer(b3b200_world* w, int which, void** devicePtr);

/* host-side shape / body tables (b3GpuNarrowPhase::getCollidablesCpu, getInternalData(),
 * getLocalSpaceAabb: b3GpuNarrowPhase.cpp:676-749, 810-813).  Copies up to capacity
 * elements into dst (may be NULL to query) and returns the element count in *count. */
#define B3B200_TBL_COLLIDABLES 0  /* b3b200_collidable */
#define B3B200_TBL_LOCAL_AABBS 1  /* b3b200_aabb, one per collidable */
#define B3B200_TBL_CONVEX 2       /* b3b200_convex_polyhedron */
#define B3B200_TBL_VERTICES 3     /* b3b200_float4 */
#define B3B200_TBL_UNIQUE_EDGES 4 /* b3b200_float4 */
#define B3B200_TBL_FACES 5        /* b3b200_face */
#define B3B200_TBL_INDICES 6      /* int */
#define B3B200_TBL_CHILD_SHAPES 7 /* b3b200_child_shape */
#define B3B200_TBL_BVH_INFOS 8    /* b3b200_bvh_info */
#define B3B200_TBL_BVH_NODES 9    /* b3b200_bvh_node */
#define B3B200_TBL_BVH_SUBTREES 10 /* b3b200_bvh_subtree */
#define B3B200_TBL_BODIES 11      /* b3b200_rigid_body (as registered; after b3b200_write_bodies: the current device state) */
#define B3B200_TBL_INERTIAS 12    /* b3b200_inertia */
int b3b200_get_table(b3b200_world* w, int which, void* dst, int capacity, int* count);

/* ------------------------------------------------- stand-alone broadphase */
/* b3GpuBroadphaseInterface (b3GpuBroadphaseInterface.h:12-40) as used by PairBench.cpp:208-379 */
int b3b200_bp_create(int kind, int device, void* stream, int maxProxies, int maxPairs, b3b200_broadphase** out);
int b3b200_bp_destroy(b3b200_broadphase* bp);
int b3b200_bp_create_proxy(b3b200_broadphase* bp, const float* aabbMin3, const float* aabbMax3, int userPtr);
int b3b200_bp_create_large_proxy(b3b200_broadphase* bp, const float* aabbMin3, const float* aabbMax3, int userPtr);
int b3b200_bp_write_aabbs(b3b200_broadphase* bp);                       /* writeAabbsToGpu */
int b3b200_bp_set_aabbs(b3b200_broadphase* bp, const b3b200_aabb* aabbs, int n); /* overwrite all AABBs (host) */
int b3b200_bp_calculate_pairs(b3b200_broadphase* bp, int maxPairs);     /* calculateOverlappingPairs */
int b3b200_bp_num_overlap(b3b200_broadphase* bp);                       /* getNumOverlap */
int b3b200_bp_get_pairs(b3b200_broadphase* bp, b3b200_int4* dst, int capacity, int* numPairs);
int b3b200_bp_device_pairs(b3b200_broadphase* bp, void** devicePtr);    /* getOverlappingPairBuffer */
int b3b200_bp_device_aabbs(b3b200_broadphase* bp, void** devicePtr);    /* getAabbBufferWS */
int b3b200_bp_last_ms(b3b200_broadphase* bp, float* ms);

/* ------------------------------------------------- multi-GPU: spatial slab decomposition (no reference counterpart)
 * A rank owns bodies [0, numOwned) of its world; the trailing bodies are ghost slots mirroring the neighbours'
 * boundary bodies.  pack selects the owned dynamic bodies whose world AABB reaches into [lo, hi] along `axis` and writes
 * b3b200_halo_record_size()-byte records (pose, velocity, inverse inertias, collidable, global id) into a DEVICE buffer;
 * the caller ships it to the neighbour (NCCL send/recv); unpack scatters received records into the ghost slots
 * [firstGhostSlot, firstGhostSlot + numGhostSlots) and parks the unused ones.  See bullet3_b200/slab.py. */
/* The same exchange driven entirely from C / C++: NCCL send/recv is called by the library on the world's stream (NCCL is
 * dlopen'ed: libnccl.so.2), one fixed-capacity message per side whose header carries the record count written on the device, so
 * a step needs no host synchronisation.  Rank 0 makes the id (b3b200_slab_unique_id), every rank gets the same bytes by its own
 * means (MPI, a file, torch.distributed), builds its world (owned bodies first, then ghost slots), uploads and calls
 * b3b200_slab_init once, b3b200_slab_exchange once, then b3b200_slab_step per step. */
int b3b200_slab_unique_id(b3b200_nccl_id* out);
int b3b200_slab_init(b3b200_world* w, const b3b200_slab_config* cfg, const b3b200_nccl_id* id);
int b3b200_slab_exchange(b3b200_world* w);
int b3b200_slab_step(b3b200_world* w, float dt);            /* b3b200_step + halo exchange, all queued on the world's stream */
int b3b200_slab_step_n(b3b200_world* w, float dt, int n);
int b3b200_slab_last_counts(b3b200_world* w, int* left, int* right);
int b3b200_slab_shutdown(b3b200_world* w);
int b3b200_halo_record_size(void);
int b3b200_halo_pack(b3b200_world* w, int axis, float lo, float hi, int numOwned, int globalIdBase, int rank, void* dstDevice,
					 int capacity, int* countOut);
int b3b200_halo_unpack(b3b200_world* w, const void* srcDevice, int count, int firstGhostSlot, int numGhostSlots);
int b3b200_halo_ghost_ids(b3b200_world* w, int* dst, int n);
/* Migration.  set_ids gives every body slot its global id (n = number of bodies; -1 = none) -- pack then sends these.
 * emigrate moves out the owned dynamic bodies whose CENTRE along `axis` lies in [lo, hi] (a range outside the slab): their
 * records go to dstDevice, their slots are freed (parked, static, id -1) and listed in slotsOut (host, `capacity` ints).
 * adopt writes `count` received records into the free slots the caller lists (host array). */
int b3b200_halo_set_ids(b3b200_world* w, const int* ids, int n);
int b3b200_halo_emigrate(b3b200_world* w, int axis, float lo, float hi, int numOwned, int rank, void* dstDevice, int capacity, int* slotsOut,
						 int* countOut);
int b3b200_halo_adopt(b3b200_world* w, const void* srcDevice, int count, const int* slots);

/* b3GpuPgsContactSolver::solveContacts / b3GpuJacobiContactSolver::solveContacts(numBodies, bodyBuf, inertiaBuf, numContacts,
 * contactBuf, config, static0Index) (b3GpuPgsContactSolver.h:34, b3GpuJacobiContactSolver.h:47) on caller-owned buffers
 * with the reference AoS layouts (device or host pointers).  `w` is the scratch context: an uploaded world with at least
 * numBodies body slots, whose solver kind / iteration count apply and whose own body state is overwritten.  On return
 * the velocities in `bodies` are the solved ones. */
int b3b200_solve_contacts_device(b3b200_world* w, int numBodies, void* bodies, const void* inertias, int numContacts, const void* contacts,
								 int static0Index);

/* ---- the callers' data formats either side of the step ----
 * Wavefront .obj -> trimesh collidable the way ConcaveScene::createConcaveMesh feeds registerConcaveMesh
 * (examples/OpenCL/rigidbody/ConcaveScene.cpp:28-109, 111-158): one fresh vertex per face corner, polygons as fans,
 * vertex = (p + shift) * scaling.  shift3 / scaling3 may be NULL.  Returns the collidable index or -1. */
int b3b200_register_concave_obj(b3b200_world* w, const char* path, const float* shift3, const float* scaling3);
/* body buffer + inertias + joints to / from a file.  load needs a world with the same shapes and body count. */
int b3b200_checkpoint_save(b3b200_world* w, const char* path);
int b3b200_checkpoint_load(b3b200_world* w, const char* path);
/* copyTransformsToVBOKernel (examples/OpenCL/rigidbody/GpuRigidBodyDemo.cpp:52-60): dst[i] = (pos.xyz, 1),
 * dst[i + numNodes] = orientation, for bodies [0, numNodes); dstDevice is a DEVICE pointer of 2 * numNodes float4
 * (e.g. a mapped graphics-interop buffer); asynchronous on the world's stream. */
int b3b200_copy_transforms(b3b200_world* w, void* dstDevice, int numNodes);

/* plain device -> host copy of a buffer obtained from b3b200_device_buffer / b3b200_bp_device_* (synchronous) */
int b3b200_device_to_host(void* dstHost, const void* srcDevice, unsigned long long bytes, int device);

/* ------------------------------------------------- parallel primitives */
/* b3RadixSort32CL::execute (b3RadixSort32CL.cpp:12-646): stable LSD sort of
 * b3SortData{key,value} by key; b3PrefixScanCL::execute (b3PrefixScanCL.cpp:45-119):
 * exclusive u32 scan, returns the total in *sum;
 * b3BoundSearchCL::execute COUNT (b3BoundSearchCL.cpp:74-203); b3FillCL (b3FillCL.cpp:41-119). */
int b3b200_radix_sort_kv(int device, b3b200_sort_data* data, int n);
int b3b200_radix_sort_keys(int device, unsigned int* keys, int n);
int b3b200_prefix_scan_u32(int device, const unsigned int* src, unsigned int* dst, int n, unsigned int* sum);
int b3b200_bound_search_count(int device, const b3b200_sort_data* sorted, int n, unsigned int* counts, int numBuckets);
int b3b200_fill_u32(int device, unsigned int* dst, unsigned int value, int n, int offset);
/* b3BoundSearchCL::execute, option BOUND_LOWER 0 / BOUND_UPPER 1 / COUNT 2 (b3BoundSearchCL.h:28-33, .cpp:74-137); dst is in/out */
int b3b200_bound_search(int device, const b3b200_sort_data* sorted, int n, unsigned int* dst, int numBuckets, int option);
/* b3PrefixScanFloat4CL::execute (b3PrefixScanFloat4CL.cpp:44-93): exclusive scan of xyz, sum = dst[n-1] like the reference */
int b3b200_prefix_scan_float4(int device, const b3b200_float4* src, b3b200_float4* dst, int n, b3b200_float4* sum);

#ifdef __cplusplus
}
#endif
#endif /* B3B200_H */
