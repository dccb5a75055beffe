/*
 * sf_canon.h -- canonical per-arena state record and its order-independent hash.
 *
 * Three implementations serialise an arena into this record so that they can be
 * compared bit for bit:
 *   - oracle/ref_harness (the UNMODIFIED reference, driven headless),
 *   - oracle/sf_oracle.c (plain-C restatement, test infrastructure),
 *   - strikeforce_b200 (sf_export_env / the device hash kernel).
 *
 * The record is a flat int32 stream of ELEMENTS:
 *     [kind, index, nfields, field_0 ... field_{n-1}]
 * The state hash is the wrapping 64-bit SUM of sf_canon_elem_hash() over all
 * elements, so it can be accumulated in any order (lanes of a warp on the GPU).
 *
 * What is covered (SURVEY.md section 8d "bit-exact check"): header counters, RNG
 * state, every human slot ever used this episode (dead slots stay observable through
 * bullet-owner credit, reference gameplay.hpp:578-596,615-633), live zombies, live
 * bullets, active portal slots, and every cell carrying dynamic content (flags
 * s[0] s[1] s[2] s[4] s[10], last-writer indices, dmg, portal_ind;
 * reference gameplay.hpp:237-243).  Render-only flags s[8], s[9] are excluded
 * (cleared by updmap before any reader, gameplay.hpp:489-495).
 */
#ifndef SF_CANON_H
#define SF_CANON_H

#include <stdint.h>

#ifdef __CUDACC__
#define SF_HD __host__ __device__
#else
#define SF_HD
#endif

enum {
    SF_K_HEADER = 1, /* index 0: mode level frame kills teams_kills loot chest ind            */
    SF_K_RNG    = 2, /* index 0: random[0..17], jomle & 0xFFFF                                 */
    SF_K_HUMAN  = 3, /* index slot: active rnpc team way f r c Hp mindamage stamina kills
                        damage effect vec ind cons[4] throw_cnt[4] blocks portals portal_ind
                        mindamage_def                                                           */
    SF_K_ZOMBIE = 4, /* index slot (live only): super f r c Hp mindamage                       */
    SF_K_BULLET = 5, /* index slot (live only): f r c df dr dc way range damage effect owner   */
    SF_K_PORTAL = 6, /* index slot (active only): f r c                                        */
    SF_K_CELL   = 7  /* index (f*N+r)*M+c (dynamic cells only): s0 hidx s1 zidx s2 bidx s4
                        ctype built(0 none 1 block 2 entrance 3 exit) dmg portal_ind           */
};

#define SF_NF_HEADER 8
#define SF_NF_RNG    19
#define SF_NF_HUMAN  27
#define SF_NF_ZOMBIE 6
#define SF_NF_BULLET 11
#define SF_NF_PORTAL 3
#define SF_NF_CELL   11

/* terminal status of an arena (0 while running) */
enum {
    SF_RUNNING   = 0,
    SF_WIN       = 1, /* Solo / Squad / Timer victory, gameplay.hpp:1145-1226            */
    SF_DEAD      = 2, /* main player Hp <= 0, gameplay.hpp:1131-1143                      */
    SF_TIMEOUT   = 3, /* Timer mode clock ran out without the kills, gameplay.hpp:1146-1153 */
    SF_TRUNCATED = 4, /* max_steps reached (new behaviour, no reference equivalent)        */
    SF_OVERFLOW  = 5, /* a slot index >= the configured capacity was needed               */
    SF_UB_GUARD  = 6  /* the reference would index themap out of bounds (update_bull,
                         gameplay.hpp:1069/1085); defined here as terminal                 */
};

/* game modes */
enum { SF_MODE_SOLO = 0, SF_MODE_TIMER = 1, SF_MODE_SQUAD = 2,
       SF_MODE_ROYALE = 3 /* "Battle Royal" / "AI Battle Royal", the online modes, gameplay.hpp:1235 */ };
#define SF_MAX_PLAYERS 16 /* players of a Battle Royale arena (BASELINE.json configs[4]) */

SF_HD static inline uint64_t sf_mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

SF_HD static inline uint64_t sf_canon_elem_begin(int kind, int index)
{
    return sf_mix64(((uint64_t)(uint32_t)kind << 32) ^ (uint64_t)(uint32_t)index ^ 0x9e3779b97f4a7c15ULL);
}

SF_HD static inline uint64_t sf_canon_elem_field(uint64_t h, int32_t v)
{
    return sf_mix64(h ^ ((uint64_t)(uint32_t)v + 0x632be59bd9b4e019ULL));
}

/* hash of a serialised record (host side) */
static inline uint64_t sf_canon_hash(const int32_t *rec, long n)
{
    uint64_t sum = 0;
    long i = 0;
    while (i + 3 <= n) {
        int kind = rec[i], index = rec[i + 1], nf = rec[i + 2];
        uint64_t h = sf_canon_elem_begin(kind, index);
        for (int k = 0; k < nf; ++k)
            h = sf_canon_elem_field(h, rec[i + 3 + k]);
        sum += h;
        i += 3 + nf;
    }
    return sum;
}

#endif /* SF_CANON_H */
