// same as src/Bullet3OpenCL/BroadphaseCollision/b3SapAabb.h:8-10
#ifndef B3_SAP_AABB_H
#define B3_SAP_AABB_H
#include "Bullet3Common/b3Scalar.h"
#include "Bullet3Collision/BroadPhaseCollision/shared/b3Aabb.h"
B3_ATTRIBUTE_ALIGNED16(struct)
b3SapAabb : public b3Aabb{};
#endif
