// C entry point over the reference's own scene generator (Source/SceneManager.cpp, compiled
// unmodified from /root/reference by oracle/Makefile into oracle/_ref/).  TEST INFRASTRUCTURE.
#include "SceneManager.h"
#include <cstdint>
#include <cstring>
extern "C" uint64_t ref_scene_generate(float particleRadius, int scene, float* pos_xyz, uint64_t cap)
{
    auto params = std::make_shared<SPHParameters<float> >();
    params->scene          = scene;
    params->particleRadius = particleRadius;
    SceneManager mgr(params);
    Vec_Vec3<float> particles, velocity;
    mgr.setupScene(particles, velocity);
    uint64_t n = particles.size();
    if(pos_xyz) std::memcpy(pos_xyz, particles.data(), sizeof(float) * 3 * (n < cap ? n : cap));
    for(auto& v : velocity) if(v[0] != 0.0f || v[1] != 0.0f || v[2] != 0.0f) return ~uint64_t(0);
    return n;
}
