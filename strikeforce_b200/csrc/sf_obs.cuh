/*
 * sf_obs.cuh -- the agent observation of bots/bot-X/Custom.hpp on the device layout.
 *
 * describe() (reference bots/bot-0.5/Custom.hpp:29-135) yields 32 floats per cell; every one
 * of them is an integer count, or an integer divided by 1000.0 or by 100.0 and rounded to
 * float.  sf_describe_milli() therefore returns the 32 features as exact integers in
 * thousandths ("milli"), and the transform of gameplay::bot() (:154-157),
 *     obs = float(pow(double(fabsf(x) / 10), 0.2)),
 * becomes one lookup pow_lut[|milli|] in a table built on the host with the host's libm --
 * the libm the reference's own bot() uses -- so observations match bit for bit.
 */
#ifndef SF_OBS_CUH
#define SF_OBS_CUH

#include "sf_core.cuh"

#define SF_OBS_R (SF_OBS_WIN / 2)
#define SF_OBS_CELLS (SF_OBS_WIN * SF_OBS_WIN)

/* Human::get_damage_effect, Character.hpp:429-443 */
SF_FN void sf_damage_effect(const SfDev &d, const SfConst &k, int env, const SfEnv &e, int h, int *dmg, int *eff)
{
    uint32_t sel = SF_AT(d.h_sel, h);
    int vec = (int)((sel >> HS_VEC_SHIFT) & 3u) - 1, ind = (int)((sel >> HS_IND_SHIFT) & 15u) - 1;
    int md = SF_AT(d.h_mind, h), st = SF_AT(d.h_stam, h);
    int pb = sf_punch_base(k, e, h);
    int base = pb > md ? pb : md;
    const SfTemplate &tp = sf_tmpl(k, h);
    if (vec == 1) {
        const SfWpn w = tp.thr[ind];
        if (0 <= st + w.stamina) {
            int a = w.damage > w.damage + md ? w.damage : w.damage + md;
            *dmg = a > base ? a : base;
            *eff = w.effect;
            return;
        }
    }
    if (vec == 2) {
        const SfWpn w = tp.wpn[ind];
        if (0 <= st + w.stamina) {
            int a = tp.shot_base[ind] > w.damage + md ? tp.shot_base[ind] : w.damage + md;
            *dmg = a > base ? a : base;
            *eff = w.effect;
            return;
        }
    }
    *dmg = base;
    *eff = 0;
}

/* describe(cell, player), Custom.hpp:29-135, in thousandths, from the cell's static byte `st` and
 * overlay word `g`.  bidx / tq: the owning bullet of the cell and its player-built record if the
 * caller already knows them, else -2 to look them up here.  With g == 0 and bidx == -1 (a cell
 * that holds nothing dynamic) this touches no memory at all. */
SF_FN void sf_describe_cell(const SfDev &d, const SfConst &k, int env, const SfEnv &e, int cell, uint32_t st, uint32_t g,
                            uint32_t viewer_team, int bidx, int tq, int32_t f[32])
{
    SF_UNROLL
    for (int i = 0; i < 32; ++i) f[i] = 0;
    uint32_t kind = (g >> C_KIND_SHIFT) & 7u;
    if (bidx == -2) bidx = sf_owner_scan(d, env, e, cell, -1);
    bool s0 = g & C_S0, s1 = g & C_S1, s2 = bidx >= 0; /* the bullet flag lives in the bullet list */
    bool s3 = (st & M_WALL) || kind == K_BLOCK;
    bool s4 = kind >= K_CHEST0 && kind < K_BLOCK;
    bool s5 = (st & M_UP) || kind == K_ENTRANCE, s6 = st & M_DOWN;
    bool s7 = (st & M_EXIT) || kind == K_EXIT;
    bool s10 = kind >= K_BLOCK;
    int occ = (int)(g & C_OCC);
    f[0] = (s0 || s1) ? 1000 : 0;
    f[1] = s2 ? 1000 : 0, f[2] = s3 ? 1000 : 0, f[3] = s4 ? 1000 : 0;
    f[4] = (s5 || s6) ? 1000 : 0;
    f[5] = s7 ? 1000 : 0, f[6] = s10 ? 1000 : 0;
    if (s0) {
        uint32_t tm = SF_AT(d.h_sel, occ) & HS_TEAM;
        if (!tm) f[9] = 1000;
        else if (tm == viewer_team) f[7] = 1000;
        else f[8] = 1000;
        uint32_t bp = SF_AT(d.h_bp, occ);
        f[11] = 1000 * SF_AT(d.h_kills, occ);
        f[12] = 1000 * (int32_t)(bp & 0xFFu);
        f[13] = 1000 * (int32_t)((bp >> 8) & 0xFFu);
        f[14] = ((bp >> 16) & 0xFFu) ? 1000 : 0;
    }
    if (s1) f[10] = 1000;
    if (s3 || s5 || s6 || s0 || s1) {
        f[15] = f[16] = 1000;
        f[17] = (s10 || s0 || s1) ? 1000 : 0;
        if (s0) f[18] = SF_AT(d.h_hp, occ);
        else if (s1) f[18] = SF_AT(d.z_hp, occ);
        else if (s10) {
            if (tq == -2) tq = sf_find_built(d, env, e, cell);
            int dmg = tq >= 0 ? SF_T(d.t_dmg, tq) : 0;
            f[18] = (s3 ? 1100 : 1000) - dmg; /* lim_block / lim_portal, gameplay.hpp:37 */
        }
    } else if (s7)
        f[15] = 1000;
    if (s0) {
        int dmg, eff;
        {   /* static subscripts only: f[] stays in registers */
            const int wy = (int)(SF_AT(d.h_pw, occ) >> POS_HI_SHIFT);
            f[20] = wy == 0 ? 1000 : 0, f[21] = wy == 1 ? 1000 : 0, f[22] = wy == 2 ? 1000 : 0, f[23] = wy == 3 ? 1000 : 0;
        }
        sf_damage_effect(d, k, env, e, occ, &dmg, &eff);
        f[24] = dmg, f[25] = -eff, f[26] = SF_AT(d.h_stam, occ);
    } else if (s1) {
        f[20] = f[21] = f[22] = f[23] = 10; /* 0.01 */
        f[24] = SF_AT(d.z_mind, occ);
    } else if (s2) {
        f[19] = 1000;
        {
            uint32_t meta = SF_AT(d.b_meta, bidx);
            int range = (int)(meta & 0xFFu), trav = (int)((meta >> 8) & 0xFFu);
            const int wy = (int)(SF_AT(d.b_pw, bidx) >> POS_HI_SHIFT), left = 10 * (range - trav); /* (range - dist) / 100.0 */
            f[20] = wy == 0 ? left : 0, f[21] = wy == 1 ? left : 0, f[22] = wy == 2 ? left : 0, f[23] = wy == 3 ? left : 0;
            f[24] = SF_AT(d.b_dmg, bidx), f[25] = -SF_AT(d.b_eff, bidx);
        }
    } else if (s7) {
        f[24] = 20, f[25] = 10;
    }
    if (s4) {
        const sf_consumable c = k.cons[(int)kind - K_CHEST0];
        f[27] = c.stamina, f[28] = c.effect, f[29] = c.hp;
    }
    if (s0) f[30] = SF_AT(d.h_dmg, occ), f[31] = -SF_AT(d.h_eff, occ);
}

/* describe() of map cell `cell`; cell < 0 is the all-zero cell `nd` used outside the map
 * (Custom.hpp:147-148) */
SF_FN void sf_describe_milli(const SfDev &d, const SfConst &k, const SfTabs &t, int env, const SfEnv &e, int cell,
                             uint32_t viewer_team, int bidx, int tq, int32_t f[32])
{
    if (cell < 0) {
        SF_UNROLL
        for (int i = 0; i < 32; ++i) f[i] = 0;
        return;
    }
    sf_describe_cell(d, k, env, e, cell, t.smap[cell], SF_G(cell), viewer_team, bidx, tq, f);
}

/* beyond the host-built table: the device's pow (counted by the caller; it may differ from
 * the host libm in the last bit).  Out of line on purpose: inlined 32 times it evicts the
 * whole observation kernel from the instruction cache. */
#ifdef __CUDACC__
__device__ __noinline__ float sf_obs_transform_slow(uint32_t a)
{
    float x = (float)((double)a / 1000.0);
    return (float)pow((double)(x / 10.0f), 0.2);
}
#else
static float sf_obs_transform_slow(uint32_t a)
{
    float x = (float)((double)a / 1000.0);
    return (float)__builtin_pow((double)(x / 10.0f), 0.2);
}
#endif

/* float(pow(double(float(m / 1000.0) / 10), 0.2)): host-built table, device pow beyond it */
SF_FN float sf_obs_transform(const SfDev &d, int32_t m, uint32_t *fallbacks)
{
    uint32_t a = m < 0 ? (uint32_t)(-(int64_t)m) : (uint32_t)m;
    if (a < (uint32_t)d.pow_lut_len) return d.pow_lut[a];
    *fallbacks += 1;
    return sf_obs_transform_slow(a);
}

/* window cell (wi, wj) of a viewer standing on `vcell` -> map cell or -1 (Custom.hpp:144-150) */
SF_FN int sf_obs_cell(int vcell, int wi, int wj)
{
    int f, r, c;
    sf_tcell_decode(vcell, &f, &r, &c);
    r += wi - SF_OBS_R, c += wj - SF_OBS_R;
    if (r < 0 || c < 0 || SF_ROWS <= r || SF_COLS <= c) return -1;
    return sf_tcell(f, r, c);
}

#endif
