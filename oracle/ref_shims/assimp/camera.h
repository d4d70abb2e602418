// TEST INFRASTRUCTURE ONLY -- stand-in for assimp's aiCamera (see assimp/types.h here).
#pragma once
#include "types.h"
struct aiCamera {
    aiString mName;
    aiVector3D mPosition;
    aiVector3D mUp;
    aiVector3D mLookAt;
    float mHorizontalFOV;
    float mClipPlaneNear;
    float mClipPlaneFar;
    float mAspect;
    aiCamera()
        : mUp(0.f, 1.f, 0.f), mLookAt(0.f, 0.f, 1.f), mHorizontalFOV(0.25f * 3.14159265358979323846f),
          mClipPlaneNear(0.1f), mClipPlaneFar(1000.f), mAspect(0.f) {}
};
