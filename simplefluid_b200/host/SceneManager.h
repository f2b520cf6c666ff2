// SceneManager.h -- initial particle sets of the four scenes (Include/SceneManager.h:25-38,
// Source/SceneManager.cpp:21-173), produced by the library's sf_scene_generate.
#pragma once
#include "SPHSolver.h"

struct SimulationScenes {
    enum Scene { SphereDrop = 0, CubeDrop, Dambreak, DoubleDambreak }; // Include/Common.h:52-58
};

class SceneManager
{
public:
    explicit SceneManager(std::shared_ptr<SPHParameters<float>>& simParams) : m_SimParams(simParams) {}

    void setupScene(Vec_Vec3<float>& particles, Vec_Vec3<float>& velocity) { fill(m_SimParams->scene, particles, velocity); }
    void setupSceneCubeDrop(Vec_Vec3<float>& p, Vec_Vec3<float>& v) { fill(SimulationScenes::CubeDrop, p, v); }
    void setupSceneSphereDrop(Vec_Vec3<float>& p, Vec_Vec3<float>& v) { fill(SimulationScenes::SphereDrop, p, v); }
    void setupSceneDambreak(Vec_Vec3<float>& p, Vec_Vec3<float>& v) { fill(SimulationScenes::Dambreak, p, v); }
    void setupSceneDoubleDambreak(Vec_Vec3<float>& p, Vec_Vec3<float>& v) { fill(SimulationScenes::DoubleDambreak, p, v); }

private:
    void fill(int scene, Vec_Vec3<float>& particles, Vec_Vec3<float>& velocity)
    {
        uint64_t n = 0;
        sf_scene_generate(m_SimParams.get(), scene, nullptr, 0, &n);
        particles.resize(n);
        if(n) sf_scene_generate(m_SimParams.get(), scene, &particles[0].x, n, &n);
        velocity.assign(particles.size(), Vec3<float>(0.f));
    }
    std::shared_ptr<SPHParameters<float>>& m_SimParams;
};
