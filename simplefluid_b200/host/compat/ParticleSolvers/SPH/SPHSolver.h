// <ParticleSolvers/SPH/SPHSolver.h> as the reference includes it (Include/SceneManager.h:21, Include/QtSPHSolver.h:22,
// Include/Controller.h:30): SPHParameters<float> and SPHSolver<float> over libsf_b200.so.
#pragma once
#include "../../../SPHSolver.h"
