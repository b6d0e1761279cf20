// TEST INFRASTRUCTURE ONLY: forwards to the GLM-compat shim (see ssb_glm_shim.hpp).
#pragma once
#include <glm/ssb_glm_shim.hpp>
