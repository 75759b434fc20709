// declaration shim, see Jolt/shapes_decl.h
#pragma once
#include <Jolt/shapes_decl.h>
