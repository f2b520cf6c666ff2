// Headless stand-in for Banana's ParticleSolvers/SPH/SPHSolver.h: only the members that
// Source/SceneManager.cpp reads.  TEST INFRASTRUCTURE.
#pragma once
#include <Banana/TypeNames.h>
namespace Banana
{
template<class T> struct SPHParameters
{
    int scene          = 0;
    T   particleRadius = T(0);
};
}
