// mock_main.cpp — runs rts_b200::RTS<mock::Soars> on a small world and prints every response as one JSON line.
// TEST INFRASTRUCTURE (tests/test_soars_adapter.py builds and runs it on the GPU box).
//   mock_main <exact|fused|tabulated> <N> <maxRefl> <maxRefr> <pulses>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "mock_soars.h"
#include "../../include/rts_soars_adapter.hpp"

int main(int argc, char **argv)
{
    using namespace mock;
    const bool fused = argc > 1 && !strcmp(argv[1], "fused");
    Parameters::rts_vars = {argc > 2 ? (unsigned)atoi(argv[2]) : 24u, argc > 3 ? (unsigned)atoi(argv[3]) : 2u, argc > 4 ? (unsigned)atoi(argv[4]) : 0u};
    const int pulses = argc > 5 ? atoi(argv[5]) : 3;

    World world;
    Transmitter tx;
    tx.pos = Vec3(0, 0, 0); tx.rot0 = SVec3(1, 0.02, 0.01); tx.az_rate = 2.0; tx.span = {0.5, 0.3, 0.0}; tx.pulses = pulses; tx.pri = 2e-3;
    Receiver rx0, rx1;
    rx0.pos = Vec3(0, 0, 0); rx0.rot0 = SVec3(1, 0.0, 0.0); rx0.sphere = {25.0, 2.5, 2.5}; rx0.noise = 120.0;
    rx1.pos = Vec3(10, -60, 5); rx1.rot0 = SVec3(1, 1.2, 0.0); rx1.sphere = {30.0, 3.0, 3.0}; rx1.noise = 80.0;
    Target plate, ball, box;
    plate.shape = "rect"; plate.w = 1.0f; plate.h = 30.0f; plate.d = 20.0f; plate.pos0 = Vec3(100, 0, 0); plate.refl = 0.8; plate.rcs0 = 3.0;
    plate.rot0 = Rotation3{0.05, 0.0, 0.0};
    ball.shape = "sphere"; ball.subdivs = 2; ball.radius = 4.0f; ball.pos0 = Vec3(60, 12, 3); ball.vel = Vec3(-40, 25, 10); ball.refl = 0.9; ball.rcs0 = 1.5;
    box.shape = "rect"; box.w = 6.0f; box.h = 4.0f; box.d = 3.0f; box.pos0 = Vec3(70, -14, -2); box.vel = Vec3(15, 30, -5); box.refl = 0.7; box.refr = 1.6;
    box.rot0 = Rotation3{0.3, 0.1, -0.2}; box.rot_rate = Rotation3{40.0, -15.0, 25.0}; box.rotating = true; box.rcs0 = 0.8;
    world.transmitters = {&tx}; world.receivers = {&rx0, &rx1}; world.targets = {&plate, &ball, &box};

    rts_b200::Options opt;
    opt.fused = fused;
    opt.tabulated = argc > 1 && !strcmp(argv[1], "tabulated");
    try {
        rts_b200::RTS<Soars>(&world, 256, 1024, opt);
    } catch (const std::exception &e) {
        fprintf(stderr, "mock_main: %s\n", e.what());
        return 1;
    }
    for (size_t j = 0; j < world.receivers.size(); j++)
        for (const Response *r : world.receivers[j]->responses)
            for (const InterpPoint &p : r->points)
                printf("{\"rx\": %zu, \"power\": %.17g, \"time\": %.17g, \"delay\": %.17g, \"doppler\": %.17g, \"phase\": %.17g, \"noise\": %.17g}\n",
                       j, p.power, p.time, p.delay, p.doppler, p.phase, p.noise_temperature);
    return 0;
}
