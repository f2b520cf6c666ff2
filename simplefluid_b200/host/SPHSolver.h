// SPHSolver.h -- C++ host facade over the C-ABI (include/sf_b200.h), shaped like the types the reference's
// solver-facing code uses:
//   SPHParameters<float>  fields written by Controller::updateSimParams (Source/Controller.cpp:54-63) + updateParams()
//   SPHSolver<float>      makeReady() / advanceFrame() called by Simulator::doSimulation (Source/Simulator.cpp:42,49)
//   Vec3 / Vec_Vec3       layout-compatible with N x 3 packed fp32 (what FluidRenderWidget uploads, .cpp:211)
// Header-only; link against simplefluid_b200/lib/libsf_b200.so.  There is no CPU fallback: every compute call
// throws SPHError when no B200 is usable.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/sf_b200.h"

template<class T> struct Vec3 {
    T x{}, y{}, z{};
    Vec3() = default;
    explicit Vec3(T s) : x(s), y(s), z(s) {}
    Vec3(T a, T b, T c) : x(a), y(b), z(c) {}
    T&       operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
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

    // EXE@0x140016650: (re)build tables/grid/walls and take the host particle set
    void makeReady()
    {
        check(sf_set_params(m_Handle, m_SimParams.get()));
        auto& x = m_SimData->particles;
        auto& v = m_SimData->velocity;
        if(v.size() != x.size()) v.assign(x.size(), Vec3<float>(0.f)); // velocity.resize(N, 0) (A.3)
        check(sf_upload_particles(m_Handle, x.empty() ? nullptr : &x[0].x, v.empty() ? nullptr : &v[0].x, static_cast<uint32_t>(x.size())));
        if(m_BoundarySeedSet) check(sf_generate_boundary(m_Handle, m_BoundarySeed));
        check(sf_make_ready(m_Handle));
        m_HostStale = false;
    }
    // EXE@0x140016810: one substep, returns the dt advanced
    float advanceFrame()
    {
        float dt = 0.f;
        check(sf_advance_frame(m_Handle, &dt));
        m_HostStale = true;
        return dt;
    }
    // inner loop of Simulator::doSimulation (Source/Simulator.cpp:46-51) without a host round trip per substep
    float advanceFrameTime(double frameTime, uint32_t* substeps = nullptr)
    {
        float t = 0.f;
        check(sf_advance_frame_time(m_Handle, frameTime, &t, substeps));
        m_HostStale = true;
        return t;
    }
    void setBoundarySeed(uint32_t seed)
    {
        m_BoundarySeed    = seed;
        m_BoundarySeedSet = true;
    }
    sf_solver* handle() { return m_Handle; }

protected:
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
