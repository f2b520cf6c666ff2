// <Banana/TypeNames.h> as the reference includes it (Include/SceneManager.h:20, Include/QtSPHSolver.h:20): Vec3,
// Vec_Vec3 and glm::length come from the B200 host facade.
#pragma once
#include "../../SPHSolver.h"
