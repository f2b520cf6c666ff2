// QtSPHSolver.h -- the solver facade the reference's Simulator owns (Include/QtSPHSolver.h:27-36), on top of the
// B200 library.  Same three accessors; the vectors are refreshed from the device on demand.
#pragma once
#include "SPHSolver.h"

class ParticleSystemData; // the viewer's container; the solver only ever shares the "Position" array with it

class QtSPHSolver : public SPHSolver<float>
{
public:
    explicit QtSPHSolver(const std::shared_ptr<SPHParameters<float>>& simParams, int device = 0) : SPHSolver<float>(simParams, device) {}
    // signature of the reference ctor; particleData is not needed headless
    QtSPHSolver(std::shared_ptr<ParticleSystemData>&, const std::shared_ptr<SPHParameters<float>>& simParams) : SPHSolver<float>(simParams, 0) {}

    unsigned int getNumParticles()
    {
        syncHost();
        return static_cast<unsigned int>(m_SimData->particles.size());
    }
    Vec_Vec3<float>& getParticles()
    {
        syncHost();
        return m_SimData->particles;
    }
    Vec_Vec3<float>& getVelocity()
    {
        syncHost();
        return m_SimData->velocity;
    }
};
