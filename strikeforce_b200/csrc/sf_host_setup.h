/*
 * sf_host_setup.h -- host-side preparation of the constants a handle uploads once:
 * what the reference recomputes every match from its text files
 * (gameplay::setup gameplay.hpp:1231-1277, Item::download_items Item.hpp:179-188,
 * Human::build Character.hpp:650-709, Random::make_p random.hpp:33-40).
 * Pure C++ (no CUDA) so that the host-check build of tests/ can share it.
 */
#ifndef SF_HOST_SETUP_H
#define SF_HOST_SETUP_H

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "sf_state.h"

namespace sfhost {

/* Character.hpp:29-45: the largest l with x / l / l ... (z divisions) >= 1, z = 2 + floor(log2 y),
 * i.e. floor(x^(1/z)) */
inline int compute_damage(int x, int y)
{
    int z = 2;
    while (1 < y) y >>= 1, ++z;
    if (x <= 0) return 0;
    int l = 0;
    for (;;) { /* grow l while (l+1)^z <= x */
        long long p = 1;
        bool fits = true;
        for (int i = 0; i < z; ++i) {
            p *= (l + 1);
            if (p > x) {
                fits = false;
                break;
            }
        }
        if (!fits) break;
        ++l;
    }
    return l;
}

struct Tables {
    std::vector<uint16_t> exp_tab, log_tab; /* 3^k - 1 ; log_3(v) at index v - 1 */
    std::vector<uint8_t> smap;
    std::vector<float> pow_lut;
};

/* discrete exp / log tables of the multiplicative group mod 65537 (generator 3) */
inline void build_rng_tables(Tables &t)
{
    t.exp_tab.assign(65536, 0);
    t.log_tab.assign(65536, 0);
    uint64_t v = 1;
    for (uint32_t k = 0; k < 65536; ++k) {
        t.exp_tab[k] = (uint16_t)(v - 1);
        t.log_tab[v - 1] = (uint16_t)k;
        v = (v * 3) % 65537;
    }
}

/* observation transform of bots/bot-0.5/Custom.hpp:157 for the value n / 1000:
 * float(pow(double(float(n / 1000.0) / 10), 0.2)), evaluated with the HOST libm -- the same
 * libm the reference's bot() uses -- so the table is bit-exact by construction */
inline float obs_transform_milli(int n)
{
    float x = (float)(n / 1000.0);
    return (float)std::pow((double)(std::fabs(x) / 10), 0.2);
}
inline void build_pow_lut(Tables &t, int len)
{
    t.pow_lut.resize((size_t)len);
    for (int n = 0; n < len; ++n) t.pow_lut[(size_t)n] = obs_transform_milli(n);
}

/* one level-up: level_solo_up / level_timer_up / level_squad_up, Character.hpp:765-801 */
struct Leveler {
    int mindamage_def, def_blocks, def_portals;
    void up(int &lvl)
    {
        ++lvl;
        mindamage_def += 5;
        if (lvl % 2 == 1) ++def_blocks, ++def_portals;
    }
};

/* Human::build, Character.hpp:650-709 */
inline void build_template(SfTemplate &tp, const int32_t *sheet, const sf_config &cfg)
{
    std::memset(&tp, 0, sizeof tp);
    tp.hp = sheet[0], tp.mindamage = sheet[1], tp.stamina = sheet[2];
    Leveler lv{sheet[1], 8, 1};
    for (int m = 0; m < 3; ++m) {
        int k = sheet[3 + m], cur = 1;
        while (--k > 0) lv.up(cur);
    }
    tp.mindamage_def = lv.mindamage_def;
    tp.blocks = lv.def_blocks, tp.portals = lv.def_portals;
    for (int i = 0; i < 4; ++i) {
        tp.cons_packed |= (uint32_t)(sheet[11 + i] & 0xFF) << (8 * i);
        int tl = sheet[15 + 2 * i], up = tl - 1 > 0 ? tl - 1 : 0;
        tp.thr_packed |= (uint32_t)(sheet[16 + 2 * i] & 0xFF) << (8 * i);
        tp.thr[i].stamina = cfg.throwables[i].stamina;
        tp.thr[i].damage = cfg.throwables[i].damage + 50 * up; /* Weapon::upgrade, Item.hpp:105-111 */
        tp.thr[i].effect = cfg.throwables[i].effect - 50 * up;
        tp.thr[i].range = cfg.throwables[i].range;
    }
    for (int i = 0; i < 8; ++i) {
        int wl = sheet[23 + i];
        tp.wpn[i].stamina = cfg.weapons[i].stamina;
        tp.wpn[i].damage = cfg.weapons[i].damage + 50 * wl;
        tp.wpn[i].effect = cfg.weapons[i].effect - 50 * wl;
        tp.wpn[i].range = cfg.weapons[i].range;
        tp.shot_base[i] = compute_damage(tp.wpn[i].damage, tp.wpn[i].range);
        if (wl) tp.w_owned |= 1u << i;
    }
}

/* returns "" or the reason the configuration cannot be represented */
inline std::string build_const(const sf_config &cfg, SfConst &k, Tables &t)
{
    if (cfg.abi_version != SF_ABI_VERSION) return "abi_version mismatch";
    if (cfg.n_envs <= 0) return "n_envs must be positive";
    if (cfg.env_id_base < 0) return "env_id_base must be non-negative (it seeds levels and the synthetic streams)";
    if (cfg.mode < SF_MODE_SOLO || cfg.mode > SF_MODE_ROYALE) return "unknown mode";
    if (cfg.mode == SF_MODE_ROYALE) {
        if (cfg.royale_players < 2 || cfg.royale_players > SF_MAX_PLAYERS) return "royale_players must be 2..16";
        if (cfg.level_min != 1 || cfg.level_max != 1) return "Battle Royale is played at level 1 (gameplay.hpp:1641)";
        for (int i = 0; i < cfg.royale_players; ++i)
            if (cfg.royale_teams[i] < 1 || cfg.royale_teams[i] > 7) return "royale_teams must be 1..7";
        if (cfg.royale_ind < 0 || cfg.royale_ind >= cfg.royale_players) return "royale_ind must name one of the players";
    } else if (cfg.royale_ind != 0) {
        return "royale_ind is a Battle Royale setting";
    }
    if (!cfg.map_cells || !cfg.map_portal) return "map_cells / map_portal missing";
    if (cfg.level_min < 1 || cfg.level_max < cfg.level_min || cfg.level_max > SF_MAX_LEVEL) return "level range";
    if (cfg.cap_humans < 10 || cfg.cap_humans > SF_LIM_HUMANS) return "cap_humans out of range (10..64)";
    if (cfg.mode == SF_MODE_ROYALE && cfg.cap_humans < cfg.royale_players) return "cap_humans below royale_players";
    if (cfg.cap_zombies < 1 || cfg.cap_zombies > SF_LIM_ZOMBIES) return "cap_zombies out of range (1..128)";
    if (cfg.cap_bullets < 1 || cfg.cap_bullets > SF_LIM_BULLETS) return "cap_bullets out of range (1..128)";
    if (cfg.cap_portals < 1 || cfg.cap_portals > SF_LIM_PORTALS) return "cap_portals out of range (1..128)";
    if (cfg.cap_built < 1 || cfg.cap_built > SF_LIM_BUILT) return "cap_built out of range (1..1023)";
    if (cfg.cap_chests < 1) return "cap_chests must be positive";
    std::memset(&k, 0, sizeof k);
    k.mode = cfg.mode, k.squad_agents = cfg.squad_agents != 0, k.auto_reset = cfg.auto_reset != 0;
    k.max_steps = cfg.max_steps, k.level_min = cfg.level_min, k.level_span = cfg.level_max - cfg.level_min + 1;
    k.n_players = cfg.mode == SF_MODE_ROYALE ? cfg.royale_players : 1;
    k.ind = cfg.mode == SF_MODE_ROYALE ? cfg.royale_ind : 0;
    k.n_agents = cfg.mode == SF_MODE_ROYALE ? k.n_players : (cfg.mode == SF_MODE_SQUAD && cfg.squad_agents) ? 10 : 1;
    for (int i = 0; i < k.n_players && cfg.mode == SF_MODE_ROYALE; ++i) k.teams[i] = (uint8_t)cfg.royale_teams[i];
    k.cap_h = cfg.cap_humans, k.cap_z = cfg.cap_zombies, k.cap_b = cfg.cap_bullets, k.cap_chest = cfg.cap_chests;
    k.cap_t = cfg.cap_built, k.cap_p = cfg.cap_portals;
    k.env_id_base = cfg.env_id_base;
    /* static map, gameplay.hpp:1252-1274: exits get portal indices in scan order */
    t.smap.assign(SF_TCELLS, M_WALL); /* tiled cell ids (sf_state.h); padding cells read as walls */
    int n_exit = 0;
    for (int lin = 0; lin < SF_CELLS; ++lin) {
        const int id = sf_tcell(lin / (SF_ROWS * SF_COLS), (lin / SF_COLS) % SF_ROWS, lin % SF_COLS);
        char c = (char)cfg.map_cells[lin];
        uint8_t m = 0;
        if (c == '#') m = M_WALL;
        else if (c == '^' || c == 'v') {
            int tgt = cfg.map_portal[lin];
            if (tgt < 0 || tgt >= SF_MAX_STATIC_EXITS) return "portal entrance target out of range";
            m = (uint8_t)((c == '^' ? M_UP : M_DOWN) | (tgt << M_TARGET_SHIFT));
        } else if (c == 'O') {
            if (n_exit >= SF_MAX_STATIC_EXITS) return "too many static exits (max 16)";
            m = M_EXIT;
            k.static_exit_cell[n_exit++] = (uint16_t)id;
        } else if (c != '.')
            return "unknown map symbol";
        t.smap[id] = m;
    }
    k.n_static_exits = n_exit;
    if (n_exit >= cfg.cap_portals) return "cap_portals must exceed the number of static exits";
    for (int id = 0; id < SF_TCELLS; ++id)
        if ((t.smap[id] & (M_UP | M_DOWN)) && (t.smap[id] >> M_TARGET_SHIFT) >= n_exit)
            return "portal entrance without an exit";
    /* the engine reads the four neighbours of zombies and the cell ahead of bullets without a
       bounds test (gameplay.hpp:664, 682, 1069): the border must not be walkable */
    for (int f = 0; f < SF_FLOORS; ++f)
        for (int r = 0; r < SF_ROWS; ++r)
            for (int c = 0; c < SF_COLS; ++c)
                if (r == 0 || c == 0 || r == SF_ROWS - 1 || c == SF_COLS - 1) {
                    uint8_t m = t.smap[sf_tcell(f, r, c)];
                    if (!(m & (M_WALL | M_UP | M_DOWN))) return "arena border must be closed";
                }
    for (int i = 0; i < 4; ++i) k.cons[i] = cfg.consumables[i];
    for (int i = 0; i < 4; ++i)
        if (cfg.throwables[i].range < 1 || cfg.throwables[i].range > 255) return "throwable range must be 1..255";
    for (int i = 0; i < 8; ++i)
        if (cfg.weapons[i].range < 1 || cfg.weapons[i].range > 255) return "weapon range must be 1..255";
    auto sheet_error = [](const int32_t *sh) -> const char * {
        for (int i = 11; i < 15; ++i)
            if (sh[i] < 0 || sh[i] > 255) return "consumable counts must be 0..255";
        for (int i = 0; i < 4; ++i)
            if (sh[16 + 2 * i] < 0 || sh[16 + 2 * i] > 255) return "throwable counts must be 0..255";
        for (int i = 3; i < 6; ++i)
            if (sh[i] < 1 || sh[i] > 400) return "sheet levels must be 1..400";
        return nullptr;
    };
    if (const char *e = sheet_error(cfg.npc_sheet)) return e;
    for (int p = 0; p < k.n_players; ++p) {
        const int32_t *sh = p == k.ind ? cfg.player_sheet : cfg.royale_sheets[p]; /* hum[ind] = me */
        if (const char *e = sheet_error(sh)) return e;
        build_template(k.players[p], sh, cfg);
        if (k.players[p].blocks > 255 || k.players[p].portals > 255) return "block / portal allowance above 255";
        k.player_punch_base[p] = compute_damage(k.players[p].mindamage_def, 1);
    }
    build_template(k.npc, cfg.npc_sheet, cfg);
    if (k.npc.blocks > 255 || k.npc.portals > 255) return "block / portal allowance above 255";
    for (int L = 1; L <= SF_MAX_LEVEL; ++L) { /* gen_human: 3 level-ups per level, Character.hpp:883-887 */
        k.npc_mindamage_def[L] = k.npc.mindamage_def + 15 * (L - 1);
        k.npc_punch_base[L] = compute_damage(k.npc_mindamage_def[L], 1);
    }
    k.npc_mindamage_def[0] = k.npc_mindamage_def[1];
    k.npc_punch_base[0] = k.npc_punch_base[1];
    build_rng_tables(t);
    return "";
}

} // namespace sfhost
#endif
