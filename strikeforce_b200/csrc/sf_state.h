/*
 * sf_state.h -- device-resident state of the batched simulator: structure-of-arrays in HBM.
 *
 * One arena ("env") of the reference is the global state of one process
 * (gameplay.hpp:37-55 bull/zomb/hum/portal/mb/active/mz/mh and `gameplay g`, :437-1739).
 * Here E arenas live side by side.  Every per-entity field is an array indexed
 * [slot][env] (env minor), so the 32 lanes of a warp -- one lane per arena, all walking the
 * same slot -- touch 32 consecutive elements.  The only per-arena contiguous block is the
 * dynamic overlay of the map grid (`grid`, one uint16 per cell), which is addressed by
 * position, not by slot.
 *
 * What is NOT stored per arena: the static map, the item tables and the character
 * templates (shared constants, SfConst), chest lists (a chest is a cell kind), per-cell
 * pointers (see below).
 *
 * Cell overlay word (uint16), the dynamic part of `struct node` (gameplay.hpp:237-243):
 *   bits 0-7   slot of the human / zombie standing here (valid under S0 / S1; a cell never
 *              holds both: humans enter '?^v.X*' cells only, zombies '.' only, :750, :682)
 *   bit  8     s[0] human     bit 9  s[1] zombie     (bit 10: s[2], the bullet flag, is a
 *              property of the bullet list and is kept here only when an arena's flag table
 *              overflows, see "the bullet flag" in sf_core.cuh; no rule reads it from a cell word)
 *   bits 11-13 kind: 0 none, 1-4 chest of type kind-1 (s[4] + cons), 5 player-built block
 *              (s[3]+s[10]), 6 player-built entrance (s[5]+s[10]), 7 player-built exit
 *              (s[7]+s[10]).  Chests and built objects both need a '.' cell, so they never
 *              coincide (:536, :706).
 *   bits 14-15 with bits 0-7: on a player-built cell nobody stands on, a validated hint of the
 *              slot of its record in the built list (sf_built_slot).
 * The reference's bullet flag s[2] and its last-writer pointer (node::bullet, trusted only under
 * s[2]) are the BF_OWNS bit of exactly one live bullet standing in that cell ("last
 * writer"); no cell word changes when a bullet moves.
 */
#ifndef SF_STATE_H
#define SF_STATE_H

#include <stdint.h>

#include "strikeforce_b200.h"

/* Cell ids on the device are TILED: the map is cut into tiles of 4 rows x 8 columns (32 cells =
 * 64 bytes of overlay, the granule in which L2 fetches from HBM), 13 x 8 tiles per floor, so the
 * four neighbours of a cell usually lie in the same granule (a row-major grid needs three).
 * id = tile << 5 | (row & 3) << 3 | (col & 7), tile = (floor * 8 + row / 4) * 13 + col / 8.
 * Cells of the padding rows 30, 31 and columns 100..103 are never referenced. */
#define SF_TILES_X ((SF_COLS + 7) / 8) /* 13 */
#define SF_TILES_Y ((SF_ROWS + 3) / 4) /* 8 */
#define SF_TCELLS (SF_FLOORS * SF_TILES_Y * SF_TILES_X * 32) /* 9,984 */
#if SF_TCELLS > 16384
#error "cell ids are 14 bits wide (position words, sf_state.h): 3 * ceil(rows/4) * ceil(cols/8) * 32 must not exceed 16,384"
#endif
#define SF_GRID_STRIDE SF_TCELLS

#ifdef __CUDACC__
#define SF_HDI __host__ __device__ __forceinline__
#else
#define SF_HDI static inline
#endif
SF_HDI int sf_tcell(int f, int r, int c)
{
    return ((((f * SF_TILES_Y + (r >> 2)) * SF_TILES_X + (c >> 3)) << 5) | ((r & 3) << 3) | (c & 7));
}
SF_HDI void sf_tcell_decode(int t, int *f, int *r, int *c)
{
    int tile = t >> 5, trow = tile / SF_TILES_X, tcol = tile - trow * SF_TILES_X;
    *f = trow / SF_TILES_Y;
    *r = ((trow % SF_TILES_Y) << 2) | ((t >> 3) & 3);
    *c = (tcol << 3) | (t & 7);
}

/* cell overlay bits */
#define C_OCC 0x00FFu
#define C_S0 0x0100u
#define C_S1 0x0200u
#define C_S2 0x0400u /* only for the flags an arena's table had no room for (sf_core.cuh, "the bullet flag") */
#define C_KIND_SHIFT 11
#define C_KIND (7u << C_KIND_SHIFT)
enum { K_NONE = 0, K_CHEST0 = 1, K_BLOCK = 5, K_ENTRANCE = 6, K_EXIT = 7 };

/* static map byte */
#define M_WALL 0x01u
#define M_UP 0x02u   /* '^' */
#define M_DOWN 0x04u /* 'v' */
#define M_EXIT 0x08u /* 'O' */
#define M_TARGET_SHIFT 4 /* bits 4-7: exit index of a static '^' / 'v' */
#define SF_MAX_STATIC_EXITS 16

/* what node::showit() prints (gameplay.hpp:321-341) */
enum { SH_WALL, SH_HUMAN, SH_ZOMBIE, SH_UP, SH_DOWN, SH_BULLET, SH_CHEST, SH_EXIT, SH_DOT };

/* human word h_sel: team | rnpc | agent | vec+1 | ind+1 | on-entrance | team bit 2.  A team is 0
 * (NPC humans) .. 7; its bits are only ever compared, sf_team_bits / sf_team_value convert. */
#define HS_TEAM 0x0803u
#define HS_RNPC 0x0004u
#define HS_AGENT 0x0008u
#define HS_VEC_SHIFT 4 /* 2 bits, stores vec + 1 */
#define HS_IND_SHIFT 6 /* 4 bits, stores ind + 1 */
#define HS_ON_ENT 0x0400u /* stands on a player-built entrance whose exit was taken (sf_obey) */
#define HS_KEEP (0x080Fu | HS_ON_ENT) /* what a new selection leaves alone */
#define sf_team_bits(team) ((((uint32_t)(team)) & 3u) | ((((uint32_t)(team)) & 4u) << 9))
#define sf_team_value(sel) ((int32_t)((((uint32_t)(sel)) & 3u) | ((((uint32_t)(sel)) >> 9) & 4u)))
/* position words: cell id in bits 0-13, (way - 1) or `super` in bits 14-15 */
#define POS_CELL 0x3FFFu
#define POS_HI_SHIFT 14
/* bullet word b_meta: range | travelled << 8 | (owner + 1) << 16 | BF_OWNS | BF_SPILL */
#define BF_OWNS 0x01000000u
#define BF_SPILL 0x02000000u /* with BF_OWNS: the flag it owns is kept in the overlay (C_S2), not in the table */

/* hard limits of this layout (sf_create validates the configured caps against them) */
#define SF_LIM_HUMANS 64
#define SF_LIM_ZOMBIES 128
#define SF_LIM_BULLETS 128
#define SF_LIM_PORTALS 128
#define SF_LIM_BUILT 1023 /* record slots fit the ten hint bits of a cell */
#define SF_MAX_LEVEL 64

typedef struct SfWpn { int32_t stamina, damage, effect, range; } SfWpn;

/* a character as the tick sees it: Human::build (Character.hpp:650-709) already applied */
typedef struct SfTemplate {
    int32_t hp, mindamage, stamina;     /* Hp = def_Hp etc. as read from the sheet */
    int32_t blocks, portals;            /* back_tmp(), Character.hpp:157-162 */
    uint32_t cons_packed, thr_packed;   /* 4 x uint8 counts */
    SfWpn thr[4];                       /* upgraded lvl-1 times, Character.hpp:676-680 */
    SfWpn wpn[8];                       /* upgraded lvl times, :683-686 */
    int32_t shot_base[8];               /* compute_damage(wpn.damage, wpn.range), :404 */
    uint32_t w_owned;                   /* bit i: weapon level > 0 */
    int32_t mindamage_def;              /* after the sheet's own level-ups */
} SfTemplate;

/* constants shared by every arena of a handle (kernel parameter) */
typedef struct SfConst {
    int32_t mode, squad_agents, auto_reset, max_steps, level_min, level_span, n_agents;
    int32_t n_players;                      /* humans 0 .. n_players-1 carry the player sheet (1 except in Battle Royale) */
    int32_t ind;                            /* `ind`: the player whose copy of the match this is (0 except in Battle Royale) */
    uint8_t teams[SF_MAX_PLAYERS];          /* Battle Royale: team of each player */
    int32_t cap_h, cap_z, cap_b, cap_chest, cap_t, cap_p;
    int64_t env_id_base;
    int32_t n_static_exits;
    uint16_t static_exit_cell[SF_MAX_STATIC_EXITS];
    sf_consumable cons[4];
    SfTemplate players[SF_MAX_PLAYERS], npc; /* [ind] = `me`; the others only in Battle Royale */
    int32_t player_punch_base[SF_MAX_PLAYERS]; /* compute_damage(mindamage_def, 1), Character.hpp:393 */
    int32_t npc_punch_base[SF_MAX_LEVEL + 1]; /* per level: gen_human's level-ups, :883-887 */
    int32_t npc_mindamage_def[SF_MAX_LEVEL + 1];
} SfConst;

/* device arrays; E = env stride (n_envs rounded up to 32) */
typedef struct SfDev {
    int32_t n_envs, E;
    int32_t cap_t;     /* player-built records per arena (multiple of 8) */
    /* header */
    uint32_t *frame;
    int32_t *kills, *tkills, *loot, *chest;
    uint32_t *misc;    /* level | status << 8 | hw_h << 16 | fast-seed flag << 24 | rng_cst bank << 25 | exit watch << 26 */
    uint32_t *steps, *episode, *ntemp;
    uint64_t *mh, *mz, *mb, *mp; /* live masks: mh[E], mz[2][E], mb[2][E], mp[2][E] */
    /* random.hpp state in the discrete-log domain */
    uint32_t *rng_log;  /* [9][E] log_3(random[2j]) | log_3(random[2j+1]) << 16 */
    uint32_t *rng_cst;  /* [2][18][E] 2*seed[i] | (2*log_3(us[i])) << 8; bank = misc bit 25, the other
                           bank belongs to the pending stream */
    uint32_t *pend_log; /* [9][E] packed logs of the pending (next episode's) stream */
    uint32_t *pend_n;   /* [E] warm-up draws made by the pending stream | fast flag << 16 */
    uint32_t *rng_w;    /* [E] sum of the values random[10..17] */
    uint32_t *jomle;
    /* humans [cap_h][E] */
    uint16_t *h_pw, *h_sel;
    uint32_t *h_bp;    /* blocks | portals << 8 | (portal_ind + 1) << 16 */
    int32_t *h_hp, *h_mind, *h_stam, *h_kills, *h_dmg, *h_eff;
    uint32_t *h_cons, *h_thr;
    uint8_t *h_cmd;    /* command[] between get_command and obey (gameplay.hpp:43, 979-1010) */
    /* zombies [cap_z][E] */
    uint16_t *z_pos;
    int32_t *z_hp, *z_mind;
    /* bullets [cap_b][E] */
    uint16_t *b_pw;
    uint32_t *b_meta;
    int32_t *b_dmg, *b_eff;
    /* player-built cells [E][cap_t], contiguous per arena (gameplay::temp, gameplay.hpp:469) */
    uint16_t *t_cell;
    int32_t *t_dmg;
    uint8_t *t_pidx;
    /* portal exits [cap_p][E] (portal[], gameplay.hpp:51) */
    uint16_t *p_cell;
    /* cell overlay [E][SF_GRID_STRIDE] */
    uint16_t *grid;
    /* per-step results and running statistics */
    sf_step_out *out;
    sf_step_out *out_mirror; /* sf_step_host: the caller's pinned host buffer as the device sees it, or NULL */
    unsigned long long *stats; /* [SF_STAT_COUNT] */
    /* shared tables in global memory */
    const uint8_t *smap;      /* [SF_TCELLS] static map bytes, tiled like the overlay */
    const uint16_t *exp_tab;  /* [65536] 3^k mod 65537, minus one */
    const uint16_t *log_tab;  /* [65536] log_3(v) for v = index + 1 */
    const float *pow_lut;     /* observation transform, see sf_observe */
    int32_t pow_lut_len;
} SfDev;

#endif
