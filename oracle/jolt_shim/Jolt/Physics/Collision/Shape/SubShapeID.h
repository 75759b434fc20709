// declaration shim, see Jolt/Jolt.h
#pragma once
#include <Jolt/Jolt.h>
