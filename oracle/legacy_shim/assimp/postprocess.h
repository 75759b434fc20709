// oracle/legacy_shim -- TEST INFRASTRUCTURE: declarations only (see SDL2/SDL.h).
#pragma once
enum { aiProcess_CalcTangentSpace = 0x1, aiProcess_JoinIdenticalVertices = 0x2, aiProcess_Triangulate = 0x8, aiProcess_GenSmoothNormals = 0x40, aiProcess_FlipUVs = 0x800000 };
