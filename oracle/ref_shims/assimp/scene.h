// TEST INFRASTRUCTURE ONLY -- just enough of assimp's scene graph types for
// /root/reference/lib/output.h's stream operators to compile. Never populated.
#pragma once
#include "camera.h"
#include "types.h"
struct aiNode {
    aiString mName;
    unsigned mNumChildren = 0;
    aiNode** mChildren = nullptr;
};
struct aiFace {
    unsigned mNumIndices = 0;
    unsigned* mIndices = nullptr;
};
struct aiMesh {
    unsigned mNumFaces = 0;
    aiFace* mFaces = nullptr;
    aiVector3D* mVertices = nullptr;
};
