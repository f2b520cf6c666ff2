// <Banana/Data/ParticleSystemData.h> as the reference includes it (Include/FluidRenderWidget.h:22).
#pragma once
#include "../../../SPHSolver.h"
