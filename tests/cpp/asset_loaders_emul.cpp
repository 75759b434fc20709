// tests/cpp/asset_loaders_emul.cpp -- TEST INFRASTRUCTURE ONLY: the host-side readers of leisure_software_renderer_b200/csrc/asset_loaders.hpp
// (what shsb_mesh_load_obj / shsb_texture_load_png parse with before they upload) behind a C interface for the CPU tests.
#include <cstring>
#include "asset_loaders.hpp"

static shsb_loaders::ObjMesh g_mesh;
static std::vector<unsigned char> g_rgba;
static std::string g_err;

extern "C"
{
    int32_t shsld_load_obj(const char* path, uint32_t counts2[2])
    {
        g_mesh = shsb_loaders::ObjMesh{};
        g_err = shsb_loaders::load_obj(path, g_mesh);
        counts2[0] = (uint32_t)(g_mesh.positions.size() / 3); counts2[1] = (uint32_t)g_mesh.indices.size();
        return g_err.empty() ? 0 : 1;
    }
    void shsld_mesh_copy(float* pos, float* nrm, float* uv, uint32_t* idx)
    {
        std::memcpy(pos, g_mesh.positions.data(), g_mesh.positions.size() * 4); std::memcpy(nrm, g_mesh.normals.data(), g_mesh.normals.size() * 4);
        std::memcpy(uv, g_mesh.uvs.data(), g_mesh.uvs.size() * 4); std::memcpy(idx, g_mesh.indices.data(), g_mesh.indices.size() * 4);
    }
    int32_t shsld_load_png(const char* path, int32_t flip_y, int32_t wh2[2])
    {
        int w = 0, h = 0;
        g_err = shsb_loaders::load_png(path, flip_y != 0, g_rgba, w, h);
        wh2[0] = w; wh2[1] = h;
        return g_err.empty() ? 0 : 1;
    }
    void shsld_png_copy(unsigned char* rgba) { std::memcpy(rgba, g_rgba.data(), g_rgba.size()); }
    const char* shsld_error() { return g_err.c_str(); }
}
