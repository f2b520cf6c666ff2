// Simulator.h -- headless restatement of the reference's simulation driver (Include/Simulator.h:36-75,
// Source/Simulator.cpp:22-103): same public methods; Qt signals become std::function callbacks; the TBB arena
// disappears (the substep is a CUDA launch sequence).  particleChanged fires once per finished frame instead of
// once per substep (the viewer can only draw frames; the reference floods its event queue).
#pragma once
#include <atomic>
#include <functional>
#include <future>
#include "QtSPHSolver.h"
#include "SceneManager.h"

class Simulator
{
public:
    explicit Simulator(int device = 0)
    {
        m_SPHSolver    = std::make_unique<QtSPHSolver>(m_SimParams, device);
        m_SceneManager = std::make_unique<SceneManager>(m_SimParams);
    }
    ~Simulator()
    {
        stop();
        if(m_SimulationFutureObj.valid()) m_SimulationFutureObj.wait();
    }

    const std::shared_ptr<SPHParameters<float>>& getSimParams() { return m_SimParams; }
    QtSPHSolver& solver() { return *m_SPHSolver; }
    float        simTime() const { return m_SimTime; }

    bool isRunning() { return !m_bStop; }
    void stop() { m_bStop = true; }
    void reset()
    {
        m_bStop = true;
        setupScene();
    }
    void startSimulation()
    {
        m_bStop = false;
        if(m_SimulationFutureObj.valid()) m_SimulationFutureObj.wait();
        m_SimulationFutureObj = std::async(std::launch::async, [&] { doSimulation(); });
    }

    // slots
    void doSimulation()
    {
        m_SPHSolver->makeReady();
        while(m_SimTime < m_SimParams->stopTime && !m_bStop) {
            // while(frameTime < 0.0333333333) frameTime += advanceFrame();  -- evaluated on the device
            const float frameTime = m_SPHSolver->advanceFrameTime(0.0333333333);
            if(particleChanged) particleChanged();
            m_SimTime += frameTime;
            if(systemTimeChanged) systemTimeChanged(m_SimTime);
            if(frameFinished) frameFinished();
        }
        if(!m_bStop) {
            m_bStop = true;
            if(simulationFinished) simulationFinished();
        }
    }
    void changeScene(SimulationScenes::Scene scene)
    {
        m_SimParams->scene = scene;
        setupScene();
    }
    void setupScene()
    {
        m_SimTime = 0;
        if(systemTimeChanged) systemTimeChanged(m_SimTime);
        if(m_SimulationFutureObj.valid()) m_SimulationFutureObj.wait(); // never touch the scene while the worker runs
        m_SceneManager->setupScene(m_SPHSolver->getParticles(), m_SPHSolver->getVelocity());
        if(particleChanged) particleChanged();
        if(numParticleChanged) numParticleChanged(m_SPHSolver->getNumParticles());
    }

    // signals
    std::function<void()>             simulationFinished, particleChanged, frameFinished;
    std::function<void(float)>        systemTimeChanged;
    std::function<void(unsigned int)> numParticleChanged;

protected:
    std::atomic<bool>                     m_bStop{ true };
    float                                 m_SimTime = 0;
    std::shared_ptr<SPHParameters<float>> m_SimParams = std::make_shared<SPHParameters<float>>();
    std::unique_ptr<SceneManager>         m_SceneManager;
    std::unique_ptr<QtSPHSolver>          m_SPHSolver;
    std::future<void>                     m_SimulationFutureObj;
};
