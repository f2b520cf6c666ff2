// QtSPHSolver.h -- the solver facade the reference's Simulator owns (Include/QtSPHSolver.h:27-36), on top of the
// B200 library.  Same constructor, same three accessors; the vectors are refreshed from the device on demand.
// This header REPLACES the reference's Include/QtSPHSolver.h: with it in place the
// reference's Source/Simulator.cpp and Source/SceneManager.cpp compile unmodified.
#pragma once
#include "SPHSolver.h"

class QtSPHSolver : public Banana::SPHSolver<float>
{
public:
    explicit QtSPHSolver(const std::shared_ptr<Banana::SPHParameters<float>>& simParams, int device = 0) : SPHSolver<float>(simParams, device) {}
    // The reference's constructor (Include/QtSPHSolver.h:30-31, called from Include/Simulator.h:42).  In the reference
    // the solver's positions ALIAS the container's "Position" array, which the renderer uploads on every
    // particleChanged signal (Source/FluidRenderWidget.cpp:204-221).  Here the array is refreshed from the device
    // after every advanceFrame() once the renderer has created it (Source/FluidRenderWidget.cpp:347), so the viewer
    // needs no change; a headless caller passes no container (or one without a "Position" array) and pays nothing.
    QtSPHSolver(std::shared_ptr<Banana::ParticleSystemData>& particleData, const std::shared_ptr<Banana::SPHParameters<float>>& simParams)
        : SPHSolver<float>(simParams, 0), m_ParticleData(particleData)
    {
    }

    unsigned int getNumParticles()
    {
        syncHost();
        return static_cast<unsigned int>(m_SimData->particles.size());
    }
    Banana::Vec_Vec3<float>& getParticles()
    {
        syncHost();
        return m_SimData->particles;
    }
    Banana::Vec_Vec3<float>& getVelocity()
    {
        syncHost();
        return m_SimData->velocity;
    }

protected:
    void afterAdvance() override
    {
        if(!m_ParticleData || !m_ParticleData->hasArray("Position")) return;
        uint32_t n = 0;
        check(sf_num_particles(m_Handle, &n));
        auto arr = m_ParticleData->getArray("Position");
        if(arr->size() != static_cast<size_t>(n) * 12) arr->bytes.resize(static_cast<size_t>(n) * 12);
        if(n) check(sf_download_positions(m_Handle, static_cast<float*>(arr->data())));
    }
    std::shared_ptr<Banana::ParticleSystemData> m_ParticleData;
};
