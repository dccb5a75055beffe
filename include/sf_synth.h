/*
 * sf_synth.h -- the synthetic workload of SURVEY.md section 8(d): per-arena seeds and
 * per-(arena, agent) action streams.  Shared by the reference harness, the C oracle,
 * the CUDA action generator and bench.py so that every arm plays the same matches.
 */
#ifndef SF_SYNTH_H
#define SF_SYNTH_H

#include <stdint.h>

#ifdef __CUDACC__
#define SF_SYNTH_HD __host__ __device__
#else
#define SF_SYNTH_HD
#endif

/* seeds of episode k of global arena e (reference serials are 30-bit, gameplay.hpp:1746) */
SF_SYNTH_HD static inline int64_t sf_synth_tb(int64_t e) { return 1700000000LL + e; }
SF_SYNTH_HD static inline int64_t sf_synth_serial(int64_t e, int64_t k)
{
    return (123456789LL + 7919LL * e + 104729LL * k) & ((1LL << 30) - 1);
}

/* splitmix64 action stream: state_0 = 42 + e*1000003 + agent, one draw per env-step */
SF_SYNTH_HD static inline uint64_t sf_synth_stream_init(int64_t e, int agent)
{
    return (uint64_t)(42LL + e * 1000003LL + agent);
}
SF_SYNTH_HD static inline uint64_t sf_synth_stream_next(uint64_t *state)
{
    uint64_t z = (*state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* action of (arena e, agent a) at global step t (t counts sf_step calls since creation) */
SF_SYNTH_HD static inline uint64_t sf_synth_draw_at(int64_t e, int agent, uint64_t t)
{
    uint64_t s = sf_synth_stream_init(e, agent) + t * 0x9E3779B97F4A7C15ULL;
    return sf_synth_stream_next(&s);
}

/* action alphabets: the 9 symbols the shipped bots expose (bots/bot-0.5/Custom.hpp:162)
 * and the 28 gameplay symbols of valid_commands (gameplay.hpp:45) minus '3' */
#define SF_ACTIONS9  "+xzqeawsd"
#define SF_ACTIONS28 "+qeuzxawsdfghjkl;'cvbnm,./[]"

#endif /* SF_SYNTH_H */
