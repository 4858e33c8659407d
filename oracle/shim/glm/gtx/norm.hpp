// TEST INFRASTRUCTURE ONLY (oracle tier A): glm/gtx/norm.hpp stand-in; length2 lives in glm.hpp.
#pragma once
#include "../glm.hpp"
