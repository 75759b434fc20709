// oracle/glm_shim: inverse()/transpose() live in glm.hpp of this shim. TEST INFRASTRUCTURE ONLY.
#pragma once
#include "../glm.hpp"
