// asset_loaders.hpp -- on-disk fixtures for the path: the two readers the reference's demos use to get
// a mesh and a texture from disk, without the third-party libraries they sit on there.
//   load_obj  : resources/loaders/mesh_loader_assimp.hpp:42-101 (load_meshes_assimp) for Wavefront OBJ.  Assimp builds one vertex per
//               distinct attribute combination and keeps the file's face order; the raster path observes exactly that (triangle
//               order + per-corner attributes), so this reader does the same: one vertex per distinct (v, vt, vn) triple in order of
//               first use, faces fan-triangulated (aiProcess_Triangulate), missing normals (0, 1, 0) and uvs (0, 0) like the
//               loader's fallbacks (:71-87).  Numbers are parsed as decimal -> double -> float.
//   load_png  : resources/loaders/texture_loader_sdl.hpp:21-56 (load_texture2d_sdl_image: IMG_Load + conversion to RGBA32 + optional
//               vertical flip, default on) for non-interlaced PNG of 8 bits per sample (grey, grey + alpha, RGB, RGBA, palette with
//               tRNS; palette / grey also at 1, 2, 4 bits).  zlib does the inflate.
// Host code only; api.cu uploads the results.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include <zlib.h>

namespace shsb_loaders
{
    struct ObjMesh { std::vector<float> positions, normals, uvs; std::vector<uint32_t> indices; };

    inline bool read_file(const char* path, std::vector<unsigned char>& out)
    {
        FILE* f = std::fopen(path, "rb");
        if (!f) return false;
        std::fseek(f, 0, SEEK_END);
        const long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        if (n < 0) { std::fclose(f); return false; }
        out.resize((size_t)n);
        const size_t got = n ? std::fread(out.data(), 1, (size_t)n, f) : 0;
        std::fclose(f);
        return got == (size_t)n;
    }

    // returns an empty string on success, else what went wrong
    inline std::string load_obj(const char* path, ObjMesh& mesh)
    {
        std::vector<unsigned char> bytes;
        if (!read_file(path, bytes)) return std::string("cannot read ") + path;
        bytes.push_back('\n'); bytes.push_back(0);
        std::vector<float> v, vt, vn;
        std::map<std::tuple<long, long, long>, uint32_t> remap;
        char* p = reinterpret_cast<char*>(bytes.data());
        auto number = [](char*& s) { char* e = s; const double d = std::strtod(s, &e); s = e; return (float)d; };
        while (*p)
        {
            char* eol = std::strchr(p, '\n');
            *eol = 0;
            while (*p == ' ' || *p == '\t') ++p;
            if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) { char* s = p + 2; for (int k = 0; k < 3; ++k) v.push_back(number(s)); }
            else if (p[0] == 'v' && p[1] == 't') { char* s = p + 2; for (int k = 0; k < 2; ++k) vt.push_back(number(s)); }
            else if (p[0] == 'v' && p[1] == 'n') { char* s = p + 2; for (int k = 0; k < 3; ++k) vn.push_back(number(s)); }
            else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t'))
            {
                std::vector<uint32_t> corner;
                char* s = p + 2;
                for (;;)
                {
                    while (*s == ' ' || *s == '\t' || *s == '\r') ++s;
                    if (!*s) break;
                    long id[3] = {0, 0, 0}; // v / vt / vn, 1-based, negative = relative to the end, 0 = absent
                    for (int k = 0; k < 3; ++k)
                    {
                        char* e = s;
                        id[k] = std::strtol(s, &e, 10);
                        s = e;
                        if (*s != '/') break;
                        ++s;
                    }
                    const long nv = (long)(v.size() / 3), nt = (long)(vt.size() / 2), nn = (long)(vn.size() / 3);
                    if (id[0] < 0) id[0] += nv + 1;
                    if (id[1] < 0) id[1] += nt + 1;
                    if (id[2] < 0) id[2] += nn + 1;
                    if (id[0] < 1 || id[0] > nv || id[1] > nt || id[2] > nn) return std::string("face index out of range in ") + path;
                    const auto key = std::make_tuple(id[0], id[1], id[2]);
                    auto it = remap.find(key);
                    if (it == remap.end())
                    {
                        it = remap.emplace(key, (uint32_t)(mesh.positions.size() / 3)).first;
                        for (int k = 0; k < 3; ++k) mesh.positions.push_back(v[(size_t)(id[0] - 1) * 3 + k]);
                        if (id[1] >= 1) { mesh.uvs.push_back(vt[(size_t)(id[1] - 1) * 2]); mesh.uvs.push_back(vt[(size_t)(id[1] - 1) * 2 + 1]); }
                        else { mesh.uvs.push_back(0.0f); mesh.uvs.push_back(0.0f); }
                        if (id[2] >= 1) for (int k = 0; k < 3; ++k) mesh.normals.push_back(vn[(size_t)(id[2] - 1) * 3 + k]);
                        else { mesh.normals.push_back(0.0f); mesh.normals.push_back(1.0f); mesh.normals.push_back(0.0f); }
                    }
                    corner.push_back(it->second);
                }
                for (size_t k = 1; k + 1 < corner.size(); ++k) { mesh.indices.push_back(corner[0]); mesh.indices.push_back(corner[k]); mesh.indices.push_back(corner[k + 1]); }
            }
            p = eol + 1;
        }
        if (mesh.indices.empty()) return std::string("no faces in ") + path;
        return std::string();
    }

    inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }

    // RGBA8, row 0 = top of the image unless flip_y (the reference's default), in which case row 0 = bottom
    inline std::string load_png(const char* path, bool flip_y, std::vector<unsigned char>& rgba, int& w, int& h)
    {
        std::vector<unsigned char> f;
        if (!read_file(path, f)) return std::string("cannot read ") + path;
        static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
        if (f.size() < 33 || std::memcmp(f.data(), sig, 8) != 0) return std::string("not a PNG file: ") + path;
        size_t at = 8;
        int depth = 0, ctype = 0, interlace = 0;
        std::vector<unsigned char> idat, plte, trns;
        bool have_ihdr = false;
        while (at + 12 <= f.size())
        {
            const uint32_t len = be32(&f[at]);
            const char* type = reinterpret_cast<const char*>(&f[at + 4]);
            if (at + 12 + (size_t)len > f.size()) return std::string("truncated chunk in ") + path;
            const unsigned char* data = &f[at + 8];
            if (!std::memcmp(type, "IHDR", 4) && len >= 13) { w = (int)be32(data); h = (int)be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12]; have_ihdr = true; }
            else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
            else if (!std::memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
            else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
            else if (!std::memcmp(type, "IEND", 4)) break;
            at += 12 + (size_t)len;
        }
        if (!have_ihdr || w <= 0 || h <= 0 || w > 65536 || h > 65536) return std::string("bad IHDR in ") + path;
        if (interlace != 0) return std::string("interlaced PNG is not supported: ") + path;
        const int channels = (ctype == 0) ? 1 : (ctype == 2) ? 3 : (ctype == 3) ? 1 : (ctype == 4) ? 2 : (ctype == 6) ? 4 : 0;
        const bool small_ok = (ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4);
        if (!channels || !(depth == 8 || small_ok)) return std::string("unsupported PNG colour type / bit depth in ") + path;
        const size_t bpp = (size_t)std::max(1, channels * depth / 8), stride = ((size_t)w * channels * depth + 7) / 8;
        std::vector<unsigned char> raw((stride + 1) * (size_t)h);
        uLongf raw_len = (uLongf)raw.size();
        if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return std::string("corrupt image data in ") + path;
        std::vector<unsigned char> prev(stride, 0), cur(stride);
        rgba.assign((size_t)w * h * 4, 0);
        for (int y = 0; y < h; ++y)
        {
            const unsigned char* row = &raw[(stride + 1) * (size_t)y];
            const int filter = row[0];
            for (size_t i = 0; i < stride; ++i)
            {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
                int pred = 0;
                if (filter == 1) pred = a;
                else if (filter == 2) pred = b;
                else if (filter == 3) pred = (a + b) / 2;
                else if (filter == 4) { const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
                else if (filter != 0) return std::string("bad filter type in ") + path;
                cur[i] = (unsigned char)(row[1 + i] + pred);
            }
            unsigned char* out = &rgba[(size_t)(flip_y ? (h - 1 - y) : y) * w * 4];
            for (int x = 0; x < w; ++x)
            {
                unsigned char px[4] = {0, 0, 0, 255};
                auto sample = [&](int s) -> int { // s-th sample of the row
                    if (depth == 8) return cur[(size_t)s];
                    const int per = 8 / depth, byte = s / per, shift = (per - 1 - s % per) * depth;
                    return (cur[(size_t)byte] >> shift) & ((1 << depth) - 1);
                };
                if (ctype == 0) { const int g = sample(x) * 255 / ((1 << depth) - 1); px[0] = px[1] = px[2] = (unsigned char)g; if (trns.size() >= 2 && sample(x) == ((trns[0] << 8) | trns[1])) px[3] = 0; }
                else if (ctype == 2) { px[0] = cur[(size_t)x * 3]; px[1] = cur[(size_t)x * 3 + 1]; px[2] = cur[(size_t)x * 3 + 2];
                                       if (trns.size() >= 6 && px[0] == trns[1] && px[1] == trns[3] && px[2] == trns[5]) px[3] = 0; }
                else if (ctype == 3) { const size_t i = (size_t)sample(x); if (i * 3 + 2 >= plte.size()) return std::string("palette index out of range in ") + path;
                                       px[0] = plte[i * 3]; px[1] = plte[i * 3 + 1]; px[2] = plte[i * 3 + 2]; if (i < trns.size()) px[3] = trns[i]; }
                else if (ctype == 4) { px[0] = px[1] = px[2] = cur[(size_t)x * 2]; px[3] = cur[(size_t)x * 2 + 1]; }
                else { std::memcpy(px, &cur[(size_t)x * 4], 4); }
                std::memcpy(out + (size_t)x * 4, px, 4);
            }
            prev.swap(cur);
        }
        return std::string();
    }
}
