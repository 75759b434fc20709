// oracle/legacy_shim -- TEST INFRASTRUCTURE (see glm/gtc/noise.hpp).
#pragma once
#include <glm/gtc/noise.hpp>
