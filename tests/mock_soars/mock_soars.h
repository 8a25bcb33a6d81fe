// mock_soars.h — a stand-in for the host simulator (SOARS) that the reference plugs into.  TEST INFRASTRUCTURE.
// It offers exactly the members rs::RTS reads (/root/reference/ray_tracer.cpp:600-648,
// 806-1014, 1190-1321) with simple closed-form behaviour, so that include/rts_soars_adapter.hpp can be compiled
// and run end to end without SOARS.  tests/test_soars_adapter.py mirrors the same formulas in numpy.
#pragma once
#include <cmath>
#include <string>
#include <vector>

namespace mock {

struct u3 { unsigned x, y, z; };
struct d3 { double x, y, z; };

struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() {}
    Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};
// spherical vector as FERS/SOARS define it: azimuth = atan2(y, x), elevation = asin(z / length)
struct SVec3 {
    double length = 0, azimuth = 0, elevation = 0;
    SVec3() {}
    SVec3(double l, double a, double e) : length(l), azimuth(a), elevation(e) {}
    explicit SVec3(const Vec3 &v)
    {
        length = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
        if (length != 0) { elevation = std::asin(v.z / length); azimuth = std::atan2(v.y, v.x); }
    }
};
struct Rotation3 { double yaw = 0, pitch = 0, roll = 0; };

struct Parameters {
    static inline u3 rts_vars = {16, 2, 0};
    static inline double c_ = 299792458.0, start_ = 0.0, rate_ = 1000.0;
    static inline bool smooth_ = false;
    static u3 GetRTSVariables() { return rts_vars; }
    static double c() { return c_; }
    static double start_time() { return start_; }
    static double cw_sample_rate() { return rate_; }
    static bool interpolate_smooth() { return smooth_; }
};

struct RadarSignal {
    double carrier = 10e9, temp = 290.0;
    double GetCarrier() const { return carrier; }
    double GetTemp() const { return temp; }
};
struct TransmitterPulse { RadarSignal *wave = nullptr; double time = 0; };

struct InterpPoint {
    double power, time, delay, doppler, phase, noise_temperature;
    InterpPoint(double p, double t, double d, double dop, double ph, double nt) : power(p), time(t), delay(d), doppler(dop), phase(ph), noise_temperature(nt) {}
};
struct Transmitter;
struct Response {
    const RadarSignal *wave; const Transmitter *trans;
    std::vector<InterpPoint> points;
    Response(const RadarSignal *w, const Transmitter *t) : wave(w), trans(t) {}
    void AddInterpPoint(const InterpPoint &p) { points.push_back(p); }
};

// antenna pattern: depends on the angle between the look direction and the boresight
inline double pattern(const SVec3 &angle, const SVec3 &ref, double peak)
{
    return peak * (1.0 + 0.25 * std::cos(angle.azimuth - ref.azimuth) * std::cos(angle.elevation - ref.elevation));
}

struct Transmitter {
    Vec3 pos; SVec3 rot0; double az_rate = 0;      // boresight azimuth drifts linearly with time
    d3 span = {0.1, 0.1, 0.0};
    RadarSignal wave; int pulses = 1; double pri = 1e-3; double peak_gain = 30.0;
    int GetPulseCount() const { return pulses; }
    void GetPulse(TransmitterPulse *p, int k) { p->wave = &wave; p->time = k * pri; }
    d3 GetTxSpan() const { return span; }
    Vec3 GetPosition(double) const { return pos; }
    SVec3 GetRotation(double t) const { return SVec3(1, rot0.azimuth + az_rate * t, rot0.elevation); }
    double GetGain(const SVec3 &angle, const SVec3 &ref, double) const { return pattern(angle, ref, peak_gain); }
};

struct Receiver {
    Vec3 pos; SVec3 rot0; d3 sphere = {2.0, 2.0, 2.0}; double noise = 100.0, peak_gain = 12.0;
    std::vector<Response *> responses;
    ~Receiver() { for (Response *r : responses) delete r; }
    void SetNoiseTemperature(double t) { noise = t; }
    double GetNoiseTemperature() const { return noise; }
    d3 GetRxSphere() const { return sphere; }
    SVec3 GetRotation(double) const { return rot0; }
    Vec3 GetPosition(double) const { return pos; }
    double GetGain(const SVec3 &angle, const SVec3 &ref, double) const { return pattern(angle, ref, peak_gain); }
    void AddResponse(Response *r) { responses.push_back(r); }
};

struct Target {
    std::string shape = "rect";
    float w = 1, h = 1, d = 1, radius = 1; unsigned subdivs = 1;
    std::string v_file, n_file;
    Vec3 pos0, vel; Rotation3 rot0, rot_rate; bool rotating = false;
    double refl = 1.0, refr = 1.0, rcs0 = 1.0;
    Vec3 GetPosition(double t) const { return Vec3(pos0.x + vel.x * t, pos0.y + vel.y * t, pos0.z + vel.z * t); }
    Rotation3 GetTargetRotation(double t) const { return Rotation3{rot0.yaw + rot_rate.yaw * t, rot0.pitch + rot_rate.pitch * t, rot0.roll + rot_rate.roll * t}; }
    std::string GetShape() const { return shape; }
    void GetRect(float &w_, float &h_, float &d_) const { w_ = w; h_ = h; d_ = d; }
    void GetSphere(unsigned &s, float &r) const { s = subdivs; r = radius; }
    void GetFile(std::string &v, std::string &n) const { v = v_file; n = n_file; }
    bool GetRotating() const { return rotating; }
    double GetReflCoeff() const { return refl; }
    double GetRefrIndex() const { return refr; }
    // bistatic-angle dependent RCS (the arguments are the summed in/out angles stored in dbuf_rcs_angle)
    double GetRCS(double azi, double ele, double) const { return rcs0 * (2.0 + 0.5 * std::cos(azi) + 0.25 * std::sin(ele)); }
};

struct World {
    std::vector<Transmitter *> transmitters;
    std::vector<Receiver *> receivers;
    std::vector<Target *> targets;
};

struct Soars {
    using World = mock::World; using Transmitter = mock::Transmitter; using Receiver = mock::Receiver; using Target = mock::Target;
    using TransmitterPulse = mock::TransmitterPulse; using Parameters = mock::Parameters; using Vec3 = mock::Vec3;
    using SVec3 = mock::SVec3; using InterpPoint = mock::InterpPoint; using Response = mock::Response;
};

} // namespace mock
