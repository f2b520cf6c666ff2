// sf_sim -- headless driver: runs a scene through the Simulator facade and optionally dumps one binary frame file
// per 1/30 s frame (header: "SFF1", uint32 n, float simTime; then n x 3 fp32 positions in original particle order).
//   sf_sim --scene Dambreak --resolution 24 --stop-time 0.5 [--dump-prefix out/frame] [--seed 0] [--pause-frame K]
// --pause-frame K: stop() after frame K, then startSimulation() again -- the GUI's Stop / Start buttons; the flow
// continues where it stopped (Source/Simulator.cpp:22-30,66-69).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <string>
#include "Simulator.h"

int main(int argc, char** argv)
{
    std::string scene = "Dambreak", dump;
    float       resolution = 24.f, stopTime = 0.5f;
    uint32_t    seed = 0;
    unsigned    pauseFrame = 0;
    for(int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        if(k == "--scene") scene = argv[i + 1];
        else if(k == "--resolution") resolution = static_cast<float>(std::atof(argv[i + 1]));
        else if(k == "--stop-time") stopTime = static_cast<float>(std::atof(argv[i + 1]));
        else if(k == "--dump-prefix") dump = argv[i + 1];
        else if(k == "--seed") seed = static_cast<uint32_t>(std::atoi(argv[i + 1]));
        else if(k == "--pause-frame") pauseFrame = static_cast<unsigned>(std::atoi(argv[i + 1]));
    }
    const char* names[4] = { "SphereDrop", "CubeDrop", "Dambreak", "DoubleDambreak" };
    int         sid      = -1;
    for(int i = 0; i < 4; ++i)
        if(scene == names[i]) sid = i;
    if(sid < 0) {
        std::fprintf(stderr, "unknown scene %s\n", scene.c_str());
        return 2;
    }
    try {
        Simulator sim;
        auto      params   = sim.getSimParams();
        params->kernelRadius = 2.0f / resolution; // Controller::updateSimParams (Source/Controller.cpp:55)
        params->stopTime     = stopTime;
        params->updateParams();
        sim.solver().setBoundarySeed(seed);
        sim.changeScene(static_cast<SimulationScenes::Scene>(sid));
        unsigned frames = 0;
        sim.frameFinished = [&] {
            ++frames;
            if(pauseFrame && frames == pauseFrame) sim.stop();
            if(dump.empty()) return;
            auto&       x = sim.solver().getParticles();
            char        name[512];
            std::snprintf(name, sizeof(name), "%s.%04u.bin", dump.c_str(), frames);
            if(FILE* f = std::fopen(name, "wb")) {
                const uint32_t n = static_cast<uint32_t>(x.size());
                const float    t = sim.simTime();
                std::fwrite("SFF1", 1, 4, f);
                std::fwrite(&n, 4, 1, f);
                std::fwrite(&t, 4, 1, f);
                if(n) std::fwrite(&x[0].x, 12, n, f);
                std::fclose(f);
            }
        };
        const unsigned n  = sim.solver().getNumParticles();
        const auto     t0 = std::chrono::steady_clock::now();
        sim.startSimulation();
        while(sim.isRunning()) std::this_thread::sleep_for(std::chrono::milliseconds(2));
        if(pauseFrame && sim.simTime() < stopTime) {
            sim.startSimulation();
            while(sim.isRunning()) std::this_thread::sleep_for(std::chrono::milliseconds(2));
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        auto&        x    = sim.solver().getParticles();
        double       cx = 0, cy = 0, cz = 0;
        for(auto& p : x) {
            cx += p.x;
            cy += p.y;
            cz += p.z;
        }
        std::printf("{\"scene\": \"%s\", \"particles\": %u, \"frames\": %u, \"sim_time\": %.6f, \"wall_s\": %.3f, \"com\": [%.6f, %.6f, %.6f]}\n",
                    scene.c_str(), n, frames, sim.simTime(), secs, n ? cx / n : 0, n ? cy / n : 0, n ? cz / n : 0);
    } catch(const SPHError& e) {
        std::fprintf(stderr, "sf_sim: %s (code %d)\n", e.what(), e.code);
        return 1;
    }
    return 0;
}
