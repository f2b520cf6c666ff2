// Test glue: the reference's own Source/SceneManager.cpp, compiled UNMODIFIED against the product's host facade
// (simplefluid_b200/host/compat/), must produce the particle sets sf_scene_generate produces.
#include "SceneManager.h" // the reference's Include/SceneManager.h
#include <cstdio>
#include <cstring>
int main()
{
    const char* names[4] = { "SphereDrop", "CubeDrop", "Dambreak", "DoubleDambreak" };
    int bad = 0;
    for(int res : { 24, 40 }) {
        for(int scene = 0; scene < 4; ++scene) {
            auto params = std::make_shared<SPHParameters<float> >();
            params->scene        = scene;
            params->kernelRadius = 2.0f / static_cast<float>(res); // Source/Controller.cpp:55
            params->updateParams();
            SceneManager    mgr(params);
            Vec_Vec3<float> particles, velocity;
            mgr.setupScene(particles, velocity);
            uint64_t n = 0;
            sf_scene_generate(params.get(), scene, nullptr, 0, &n);
            std::vector<float> x(3 * n);
            sf_scene_generate(params.get(), scene, x.data(), n, &n);
            const bool same = n == particles.size() && velocity.size() == n && (n == 0 || std::memcmp(x.data(), particles.data(), 12 * n) == 0);
            std::printf("%s@%d %zu %s\n", names[scene], res, particles.size(), same ? "identical" : "DIFFERENT");
            bad += same ? 0 : 1;
        }
    }
    return bad;
}
