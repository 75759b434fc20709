// oracle/legacy_shim -- TEST INFRASTRUCTURE: declarations only (see SDL2/SDL.h).
#pragma once
struct aiVector3D { float x, y, z; };
struct aiFace { unsigned int mNumIndices; unsigned int* mIndices; };
struct aiMesh
{
    unsigned int mNumVertices, mNumFaces;
    aiVector3D* mVertices;
    aiVector3D* mNormals;
    aiVector3D* mTextureCoords[8];
    aiFace* mFaces;
    bool HasNormals() const { return mNormals != nullptr; }
    bool HasTextureCoords(unsigned int i) const { return mTextureCoords[i] != nullptr; }
};
struct aiNode;
struct aiScene { unsigned int mFlags; aiNode* mRootNode; unsigned int mNumMeshes; aiMesh** mMeshes; };
#define AI_SCENE_FLAGS_INCOMPLETE 1
