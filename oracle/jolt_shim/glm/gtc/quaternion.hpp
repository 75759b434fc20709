// empty: shs/geometry/jolt_adapter.hpp includes it but uses no quaternion in the light-culling path (see Jolt/Jolt.h)
#pragma once
#include <glm/glm.hpp>
