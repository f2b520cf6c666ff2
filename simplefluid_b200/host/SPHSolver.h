// SPHSolver.h -- C++ host facade over the C-ABI (include/sf_b200.h), shaped like the types the reference's
// solver-facing code takes from its private Banana library:
//   Vec3 / Vec_Vec3 / glm::length  what Source/SceneManager.cpp:43-169 does arithmetic with; layout-compatible with
//                                  N x 3 packed fp32 (what FluidRenderWidget uploads, Source/FluidRenderWidget.cpp:211)
//   SPHParameters<float>           fields written by Controller::updateSimParams (Source/Controller.cpp:54-63) + updateParams()
//   ParticleSystemData             the viewer's container (Source/Simulator.cpp:96-99, Source/FluidRenderWidget.cpp:211,347)
//   SPHSolver<float>               makeReady() / advanceFrame() called by Simulator::doSimulation (Source/Simulator.cpp:42,49)
// Everything lives in namespace Banana like the originals (Include/Common.h:103 does `using namespace Banana`), and the
// forwarding headers under host/compat/ give these types the include paths the reference uses
// (<Banana/TypeNames.h>, <ParticleSolvers/SPH/SPHSolver.h>, ...): the reference's own Source/SceneManager.cpp and
// Source/Simulator.cpp compile UNMODIFIED against them (tests/test_reference_callers.py).
// Header-only; link against simplefluid_b200/lib/libsf_b200.so.  There is no CPU fallback: every compute call
// throws SPHError when no B200 is usable.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/sf_b200.h"

namespace Banana
{
template<class T> struct Vec3 {
    T x{}, y{}, z{};
    Vec3() = default;
    explicit Vec3(T s) : x(s), y(s), z(s) {}
    // mixed argument types as in `Vec3<int> grid(float, float, float)` (float -> int truncation, SceneManager.cpp:46-48)
    // and `Vec3<float>(i, j, k)` with int loop counters (SceneManager.cpp:57)
    template<class A, class B, class C> Vec3(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
    T&       operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    Vec3& operator+=(const Vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    Vec3& operator-=(const Vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
};
// componentwise, one separately rounded operation each (what glm::tvec3 does)
template<class T> inline Vec3<T> operator+(const Vec3<T>& a, const Vec3<T>& b) { return Vec3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template<class T> inline Vec3<T> operator-(const Vec3<T>& a, const Vec3<T>& b) { return Vec3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template<class T> inline Vec3<T> operator*(T s, const Vec3<T>& a) { return Vec3<T>(s * a.x, s * a.y, s * a.z); }
template<class T> inline Vec3<T> operator*(const Vec3<T>& a, T s) { return Vec3<T>(a.x * s, a.y * s, a.z * s); }
template<class T> using Vec_Vec3 = std::vector<Vec3<T>>;
static_assert(sizeof(Vec3<float>) == 12, "Vec3<float> must be three packed floats");

struct SPHError : std::runtime_error {
    int code;
    SPHError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

template<class T> struct SPHParameters;
template<> struct SPHParameters<float> : sf_params {
    SPHParameters() { sf_params_default(this); }
    void updateParams() { sf_params_update(this); } // Source/Controller.cpp:63
};

// The viewer's particle container, reduced to what the solver side and FluidRenderWidget::updateParticleData touch:
// named uint properties, the particle radius and named arrays of N x C elements ("Position" is N x 3 fp32, the array
// the reference's solver aliases -- EXE@0x140011fa0 -- and the renderer uploads).
class ParticleSystemData
{
public:
    struct Array {
        std::vector<unsigned char> bytes;
        size_t                     elemBytes = 0;
        void*                      data() { return bytes.data(); }
        size_t                     size() const { return bytes.size(); } // bytes, as uploadDataAsync(data, 0, size) expects
    };
    void         setNumParticles(unsigned int n)
    {
        m_NumParticles = n;
        for(auto& kv : m_Arrays) kv.second->bytes.resize(static_cast<size_t>(n) * kv.second->elemBytes);
    }
    unsigned int getNumParticles() const { return m_NumParticles; }
    void         setUInt(const std::string& name, unsigned int v) { m_UInts[name] = v; }
    unsigned int getUInt(const std::string& name) { return m_UInts[name]; }
    void         setParticleRadius(float r) { m_Radius = r; }
    template<class T> T getParticleRadius() const { return static_cast<T>(m_Radius); }
    template<class T, int N> void addArray(const std::string& name)
    {
        auto a       = std::make_shared<Array>();
        a->elemBytes = sizeof(T) * N;
        a->bytes.resize(static_cast<size_t>(m_NumParticles) * a->elemBytes);
        m_Arrays[name] = a;
    }
    bool                   hasArray(const std::string& name) const { return m_Arrays.count(name) != 0; }
    std::shared_ptr<Array> getArray(const std::string& name) { return m_Arrays.at(name); }

private:
    unsigned int                                  m_NumParticles = 0;
    float                                         m_Radius = 0.f;
    std::map<std::string, unsigned int>           m_UInts;
    std::map<std::string, std::shared_ptr<Array>> m_Arrays;
};

// particles / velocity live on the host exactly as in the reference (SceneManager fills them by reference,
// Source/SceneManager.cpp:50-63); the device copy is authoritative between makeReady() and the next host read.
struct SPHSimData {
    Vec_Vec3<float> particles;
    Vec_Vec3<float> velocity;
};

template<class T> class SPHSolver;
template<> class SPHSolver<float>
{
public:
    explicit SPHSolver(const std::shared_ptr<SPHParameters<float>>& params, int device = 0) : m_SimParams(params)
    {
        check(sf_create(params.get(), device, &m_Handle), nullptr);
    }
    virtual ~SPHSolver() { sf_destroy(m_Handle); }
    SPHSolver(const SPHSolver&) = delete;
    SPHSolver& operator=(const SPHSolver&) = delete;

    // EXE@0x140016650: (re)build tables / grid / walls from the parameters and take the particle set.  The reference
    // re-runs it on its live vectors at every startSimulation() (Source/Simulator.cpp:42), so a stop() followed by
    // startSimulation() continues the flow: when substeps ran since the last host read, the device state IS the live
    // particle set and is kept; the host vectors are uploaded only when the host side changed them (setupScene).
    void makeReady()
    {
        check(sf_set_params(m_Handle, m_SimParams.get()));
        if(!m_HostStale) {
            auto& x = m_SimData->particles;
            auto& v = m_SimData->velocity;
            if(v.size() != x.size()) v.assign(x.size(), Vec3<float>(0.f)); // velocity.resize(N, 0) (A.3)
            check(sf_upload_particles(m_Handle, x.empty() ? nullptr : &x[0].x, v.empty() ? nullptr : &v[0].x, static_cast<uint32_t>(x.size())));
        }
        if(m_BoundarySeedSet) check(sf_generate_boundary(m_Handle, m_BoundarySeed));
        check(sf_make_ready(m_Handle));
    }
    // EXE@0x140016810: one substep, returns the dt advanced
    float advanceFrame()
    {
        float dt = 0.f;
        check(sf_advance_frame(m_Handle, &dt));
        m_HostStale = true;
        afterAdvance();
        return dt;
    }
    // inner loop of Simulator::doSimulation (Source/Simulator.cpp:46-51) without a host round trip per substep
    float advanceFrameTime(double frameTime, uint32_t* substeps = nullptr)
    {
        float t = 0.f;
        check(sf_advance_frame_time(m_Handle, frameTime, &t, substeps));
        m_HostStale = true;
        afterAdvance();
        return t;
    }
    void setBoundarySeed(uint32_t seed)
    {
        m_BoundarySeed    = seed;
        m_BoundarySeedSet = true;
    }
    sf_solver* handle() { return m_Handle; }

protected:
    virtual void afterAdvance() {}
    // refresh the host vectors from the device when a substep ran since the last read
    void syncHost()
    {
        if(!m_HostStale) return;
        uint32_t n = 0;
        check(sf_num_particles(m_Handle, &n));
        m_SimData->particles.resize(n);
        m_SimData->velocity.resize(n);
        if(n) {
            check(sf_download_positions(m_Handle, &m_SimData->particles[0].x));
            check(sf_download_velocities(m_Handle, &m_SimData->velocity[0].x));
        }
        m_HostStale = false;
    }
    void check(int rc) { check(rc, m_Handle); }
    static void check(int rc, sf_solver* h)
    {
        if(rc != SF_OK) throw SPHError(rc, sf_last_error(h));
    }

    std::shared_ptr<SPHParameters<float>> m_SimParams;
    std::unique_ptr<SPHSimData>           m_SimData = std::make_unique<SPHSimData>();
    sf_solver*                            m_Handle  = nullptr;
    bool                                  m_HostStale = false, m_BoundarySeedSet = false;
    uint32_t                              m_BoundarySeed = 0;
};
} // namespace Banana

namespace glm
{
// glm::length(vec3) = sqrt(dot(v, v)), dot = (x*x + y*y) + z*z  (Source/SceneManager.cpp:83)
inline float length(const Banana::Vec3<float>& a) { return std::sqrt((a.x * a.x + a.y * a.y) + a.z * a.z); }
} // namespace glm

#ifndef SF_B200_NO_GLOBAL_NAMES
using namespace Banana; // Include/Common.h:103
#endif
