// oracle/glm_shim: pi<T>() / half_pi<T>() live in glm.hpp of this shim. TEST INFRASTRUCTURE ONLY.
#pragma once
#include "../glm.hpp"
