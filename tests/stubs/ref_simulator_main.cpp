// Test glue: the reference's own Source/Simulator.cpp + Source/SceneManager.cpp compiled UNMODIFIED against the
// product's host facade, with the product's QtSPHSolver.h in place of Include/QtSPHSolver.h.
// Defines what moc would generate for the signals, drives the Simulator like MainWindow does and dumps the result:
//   ref_simulator <scene 0..3> <resolution> <stopTime> <out.bin> [pauseAfterFrames]
#include "Simulator.h" // the reference's Include/Simulator.h
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

static unsigned g_frames = 0, g_particleChanged = 0, g_pauseAfter = 0;
static Simulator* g_sim = nullptr;
void Simulator::simulationFinished() {}
void Simulator::systemTimeChanged(float) {}
void Simulator::numParticleChanged(unsigned int) {}
void Simulator::particleChanged() { ++g_particleChanged; }
void Simulator::frameFinished()
{
    ++g_frames;
    if(g_pauseAfter && g_frames == g_pauseAfter) g_sim->stop(); // the GUI's Stop button (Source/MainWindow.cpp)
}

int main(int argc, char** argv)
{
    if(argc < 5) return 2;
    const int   scene = std::atoi(argv[1]);
    const float res = static_cast<float>(std::atof(argv[2])), stopTime = static_cast<float>(std::atof(argv[3]));
    g_pauseAfter = argc > 5 ? static_cast<unsigned>(std::atoi(argv[5])) : 0u;
    try {
        auto particleData = std::make_shared<ParticleSystemData>();
        particleData->addArray<float, 3>("Position"); // FluidRenderWidget::initParticleDataObj (Source/FluidRenderWidget.cpp:347)
        Simulator sim(particleData);
        g_sim = &sim;
        auto params = sim.getSimParams();
        params->kernelRadius = 2.0f / res; // Controller::updateSimParams (Source/Controller.cpp:55-63)
        params->stopTime     = stopTime;
        params->updateParams();
        sim.changeScene(static_cast<SimulationScenes::Scene>(scene));
        sim.startSimulation();
        while(sim.isRunning()) std::this_thread::sleep_for(std::chrono::milliseconds(2));
        if(g_pauseAfter) { // Start again after Stop: the flow must continue where it stopped (Source/Simulator.cpp:22-30)
            sim.startSimulation();
            while(sim.isRunning()) std::this_thread::sleep_for(std::chrono::milliseconds(2));
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(20)); // the worker has left doSimulation
        // what FluidRenderWidget::updateParticleData would upload (Source/FluidRenderWidget.cpp:211)
        auto           arr = particleData->getArray("Position");
        const uint32_t n   = particleData->getNumParticles();
        if(arr->size() != static_cast<size_t>(n) * 12) return 4;
        FILE* f = std::fopen(argv[4], "wb");
        if(!f) return 5;
        std::fwrite(&n, 4, 1, f);
        std::fwrite(&g_frames, 4, 1, f);
        std::fwrite(&g_particleChanged, 4, 1, f);
        std::fwrite(arr->data(), 1, arr->size(), f);
        std::fclose(f);
        std::printf("particles %u frames %u particleChanged %u\n", n, g_frames, g_particleChanged);
    } catch(const SPHError& e) {
        std::fprintf(stderr, "ref_simulator: %s (code %d)\n", e.what(), e.code);
        return 3;
    }
    return 0;
}
