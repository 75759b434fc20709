// oracle/legacy_shim -- TEST INFRASTRUCTURE: declarations only (see SDL2/SDL.h).  Assimp is the reference's mesh importer
// (shs::ModelGeometry, shs_renderer.hpp:1248-1299); the harness feeds triangle soups directly and never loads a file.
#pragma once
struct aiScene;
namespace Assimp
{
    class Importer
    {
    public:
        const aiScene* ReadFile(const char*, unsigned int);
        const char* GetErrorString() const;
    };
}
