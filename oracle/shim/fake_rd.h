// Force-included (-include) into every reference translation unit of the oracle build.
//
// (1) The reference seeds a second engine `gen` from std::random_device
//     (reference src/blockmodel.hh:17-18) and draws the categorical proposal from it
//     (reference src/blockmodel.cc:627-628), which makes it non-reproducible even with
//     --seed.  This header swaps std::random_device for a deterministic stand-in whose
//     value the driver controls, WITHOUT editing any reference source.
// (2) With -DORACLE_LOG_RNG it also swaps std::mt19937 for a subclass that counts the
//     32-bit words drawn, so the tests can pin the draw-stream accounting
//     (engine/gen word counts).  Values are unchanged.
//
// TEST INFRASTRUCTURE (oracle build only) -- never included by product code.
#pragma once
#include <random>
#include <algorithm>
extern "C" unsigned int oracle_fake_rd_value;
namespace std {
struct fake_random_device {
    typedef unsigned int result_type;
    fake_random_device() {}
    unsigned int operator()() { return oracle_fake_rd_value; }
    static constexpr unsigned int min() { return 0u; }
    static constexpr unsigned int max() { return 0xffffffffu; }
};
#ifdef ORACLE_LOG_RNG
struct logging_mt19937 : public mt19937 {
    using mt19937::mt19937;
    unsigned long long words_drawn = 0;
    result_type operator()() { ++words_drawn; return mt19937::operator()(); }
};
#endif
}  // namespace std
#define random_device fake_random_device
#ifdef ORACLE_LOG_RNG
#define mt19937 logging_mt19937
#endif
